"""chromap_b200 — B200-native replacement for Chromap's per-read mapping hot path.

Thin ctypes view of the C ABI in include/chromap_b200.h (chromap_b200/libchromap_b200.so, built by
`__graft_entry__.build()` / chromap_b200/csrc/Makefile).  There is no CPU fallback: importing works
anywhere, but creating a Mapper without the CUDA library or without a GPU raises.
"""
from .binding import (Mapper, Params, PE_RECORD, PAIRS_RECORD, PAIR_TRACE, SAM_RECORD, Timing, CmxError, lib_path, load_library,
                      make_params, taskloop_chunks, format_sam, format_paf, exchange_finish, ExchangeStats)

__all__ = ["format_sam", "format_paf", "Mapper", "Params", "PE_RECORD", "PAIRS_RECORD", "PAIR_TRACE", "SAM_RECORD", "Timing", "CmxError", "lib_path", "load_library",
           "make_params", "taskloop_chunks", "exchange_finish", "ExchangeStats"]
