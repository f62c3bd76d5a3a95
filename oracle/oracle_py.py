"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/liboracle.so (see oracle_chromap.h).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "error_threshold", "min_num_seeds", "max_seed_freq0", "max_seed_freq1", "max_num_best_mappings",
        "max_insert_size", "mapq_threshold", "min_read_length", "drop_repetitive_reads", "trim_adapters",
        "remove_pcr_duplicates", "tn5_shift", "split_alignment", "low_memory_mode", "output_format", "single_end")]


PE_RECORD = np.dtype([("read_id", "<u4"), ("rid", "<u4"), ("fragment_start", "<u4"), ("fragment_length", "<u2"),
                      ("mapq", "u1"), ("direction", "u1"), ("is_unique", "u1"), ("num_dups", "u1"),
                      ("positive_alignment_length", "<u2"), ("negative_alignment_length", "<u2")], align=True)

PAIRS_RECORD = np.dtype([("read_id", "<u4"), ("rid1", "<u4"), ("rid2", "<u4"), ("pos1", "<u4"), ("pos2", "<u4"), ("strand1", "u1"),
                         ("strand2", "u1"), ("mapq", "u1"), ("is_unique", "u1")], align=True)

TRACE = np.dtype([("n_minimizers", "<i4", 2), ("n_pos_candidates_gen", "<i4", 2), ("n_neg_candidates_gen", "<i4", 2),
                  ("n_pos_candidates", "<i4", 2), ("n_neg_candidates", "<i4", 2), ("n_pos_mappings", "<i4", 2),
                  ("n_neg_mappings", "<i4", 2), ("min_errors", "<i4", 2), ("second_min_errors", "<i4", 2),
                  ("n_best", "<i4", 2), ("n_second_best", "<i4", 2), ("repetitive_seed_length", "<u4", 2),
                  ("supplement_result", "<i4"), ("min_sum_errors", "<i4"), ("second_min_sum_errors", "<i4"),
                  ("n_best_pairs", "<i4"), ("n_second_best_pairs", "<i4"), ("n_records", "<i4"),
                  ("trimmed_len", "<i4", 2)], align=True)


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "oracle_chromap.cc")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32, i64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int64
        L.orc_default_params.argtypes = [C.POINTER(Params)]
        L.orc_apply_preset.argtypes = [C.POINTER(Params), C.c_char_p]
        L.orc_index_load.restype = vp; L.orc_index_load.argtypes = [C.c_char_p]
        L.orc_index_build.restype = vp; L.orc_index_build.argtypes = [vp, i32, i32]
        L.orc_index_save.argtypes = [vp, C.c_char_p]
        L.orc_index_from_arrays.restype = vp; L.orc_index_from_arrays.argtypes = [i32, i32, u32, vp, vp, vp, vp, u32]
        L.orc_index_free.argtypes = [vp]
        L.orc_index_k.argtypes = [vp]; L.orc_index_w.argtypes = [vp]
        L.orc_index_arrays.restype = u32
        L.orc_index_arrays.argtypes = [vp] + [C.POINTER(vp)] * 4 + [C.POINTER(u32)]
        L.orc_index_lookup.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.orc_reference_load.restype = vp; L.orc_reference_load.argtypes = [C.c_char_p]
        L.orc_reference_from_memory.restype = vp; L.orc_reference_from_memory.argtypes = [u32, vp, vp, vp]
        L.orc_reference_free.argtypes = [vp]
        L.orc_reference_num_sequences.restype = u32; L.orc_reference_num_sequences.argtypes = [vp]
        L.orc_reference_length.restype = u32; L.orc_reference_length.argtypes = [vp, u32]
        L.orc_reference_name.restype = C.c_char_p; L.orc_reference_name.argtypes = [vp, u32]
        L.orc_reference_seq.restype = vp; L.orc_reference_seq.argtypes = [vp, u32]
        L.orc_minimizers.argtypes = [vp, u32, u32, i32, i32, vp, vp, i32]
        L.orc_banded_align.argtypes = [i32, vp, vp, i32, C.POINTER(i32)]
        L.orc_banded_traceback.argtypes = [i32, i32, vp, vp, i32, C.POINTER(i32)]
        L.orc_mapper_create.restype = vp; L.orc_mapper_create.argtypes = [C.POINTER(Params), vp, vp]
        L.orc_mapper_free.argtypes = [vp]
        L.orc_map_pairs_mt.restype = i64
        L.orc_map_pairs_mt.argtypes = [vp, u32, vp, vp, vp, vp, u32, vp, i64, i32, vp]
        L.orc_ref_task_chunks.argtypes = [u32, vp, vp, i32]
        L.orc_postprocess.restype = i64; L.orc_postprocess.argtypes = [C.POINTER(Params), vp, i64]
        L.orc_format_bed.restype = i64; L.orc_format_bed.argtypes = [vp, vp, i64, vp, i64]
        L.orc_format_tagalign.restype = i64; L.orc_format_tagalign.argtypes = [vp, vp, i64, vp, i64]
        L.orc_whitelist_load.restype = vp; L.orc_whitelist_load.argtypes = [C.c_char_p, u32]
        L.orc_whitelist_free.argtypes = [vp]
        L.orc_whitelist_sample.argtypes = [vp, vp, u64, u32, u64, u32]
        L.orc_whitelist_arrays.restype = u64; L.orc_whitelist_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
        L.orc_mapper_set_barcodes.argtypes = [vp, vp, i32, C.c_double, i32]
        L.orc_map_pairs_bc.restype = i64
        L.orc_map_pairs_bc.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, u32, u32, vp, vp, i64, i32, vp]
        L.orc_postprocess_bc.restype = i64; L.orc_postprocess_bc.argtypes = [C.POINTER(Params), vp, vp, i64]
        L.orc_format_bed_bc.restype = i64; L.orc_format_bed_bc.argtypes = [vp, vp, vp, i64, u32, vp, i64]
        L.orc_run_files_bc.argtypes = [C.POINTER(Params)] + [C.c_char_p] * 7 + [i32, vp]
        L.orc_run_files.argtypes = [C.POINTER(Params)] + [C.c_char_p] * 5 + [i32, C.POINTER(C.c_double), C.POINTER(u64)]
        L.orc_map_sam_cores.restype = i64; L.orc_map_sam_cores.argtypes = [vp, u32, vp, vp, vp, vp, u32, vp, i64]
        L.orc_run_files_sam.argtypes = [C.POINTER(Params)] + [C.c_char_p] * 5
        L.orc_run_files_paf.argtypes = [C.POINTER(Params)] + [C.c_char_p] * 5
        L.orc_map_reads_se_bc.restype = i64; L.orc_map_reads_se_bc.argtypes = [vp, u32, vp, vp, vp, vp, u32, u32, vp, vp, i64, i32, vp]
        L.orc_map_reads_se.restype = i64; L.orc_map_reads_se.argtypes = [vp, u32, vp, vp, u32, vp, i64, i32]
        L.orc_postprocess_se.restype = i64; L.orc_postprocess_se.argtypes = [C.POINTER(Params), vp, i64]
        L.orc_run_files_se.argtypes = [C.POINTER(Params)] + [C.c_char_p] * 4 + [i32]
        _lib = L
    return _lib


def make_params(preset="", **kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    if lib().orc_apply_preset(C.byref(p), preset.encode()) != 0:
        raise ValueError("unknown preset " + preset)
    for k, v in kw.items():
        setattr(p, k, int(v))
    return p


class Reference:
    def __init__(self, path=None, seqs=None):
        L = lib()
        if path is not None:
            self.h = L.orc_reference_load(path.encode())
        else:
            concat = np.concatenate(seqs).astype(np.uint8)
            offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(s) for s in seqs])
            self.h = L.orc_reference_from_memory(len(seqs), concat.ctypes.data, offs.ctypes.data, None)
        if not self.h:
            raise IOError("cannot load reference")
        self.n = L.orc_reference_num_sequences(self.h)
        self.lengths = np.array([L.orc_reference_length(self.h, i) for i in range(self.n)], dtype=np.uint32)
        self.names = [L.orc_reference_name(self.h, i).decode() for i in range(self.n)]

    def seq(self, rid):
        n = int(self.lengths[rid])
        return np.ctypeslib.as_array(C.cast(lib().orc_reference_seq(self.h, rid), C.POINTER(C.c_uint8)), (n,)).copy()


class Index:
    def __init__(self, path=None, ref=None, k=17, w=7, arrays=None):
        L = lib()
        if arrays is not None:
            a = arrays
            fl = np.ascontiguousarray(a["flags"], dtype=np.uint32); ke = np.ascontiguousarray(a["keys"], dtype=np.uint64)
            va = np.ascontiguousarray(a["vals"], dtype=np.uint64); oc = np.ascontiguousarray(a["occ"], dtype=np.uint64)
            self.h = L.orc_index_from_arrays(k, w, a["n_buckets"], fl.ctypes.data, ke.ctypes.data, va.ctypes.data,
                                             oc.ctypes.data if len(oc) else None, len(oc))
        else:
            self.h = L.orc_index_load(path.encode()) if path else L.orc_index_build(ref.h, k, w)
        if not self.h:
            raise IOError("cannot load index")
        self.k, self.w = L.orc_index_k(self.h), L.orc_index_w(self.h)

    def arrays(self):
        L = lib()
        ptrs = [C.c_void_p() for _ in range(4)]
        n_occ = C.c_uint32()
        nb = L.orc_index_arrays(self.h, *[C.byref(p) for p in ptrs], C.byref(n_occ))
        nf = 1 if nb < 16 else nb >> 4

        def arr(p, dt, n):
            if n == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n * np.dtype(dt).itemsize,)).view(dt).copy()
        return dict(n_buckets=nb, flags=arr(ptrs[0], np.uint32, nf), keys=arr(ptrs[1], np.uint64, nb),
                    vals=arr(ptrs[2], np.uint64, nb), occ=arr(ptrs[3], np.uint64, n_occ.value))

    def save(self, path):
        return lib().orc_index_save(self.h, path.encode())


def minimizers(seq, k, w, seq_index=0):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    cap = max(1, len(seq))
    h = np.zeros(cap, dtype=np.uint64)
    t = np.zeros(cap, dtype=np.uint64)
    n = lib().orc_minimizers(seq.ctypes.data, len(seq), seq_index, k, w, h.ctypes.data, t.ctypes.data, cap)
    return h[:n], t[:n]


def map_pairs(params, index, ref, seq1, off1, seq2, off2, first_read_id=0, n_threads=1, trace=False):
    """One reference batch.  seq*: uint8 concatenations, off*: uint32[n+1].  Returns (records, trace|None)."""
    L = lib()
    m = L.orc_mapper_create(C.byref(params), index.h, ref.h)
    if not m:
        raise ValueError("unsupported parameters for the oracle (BED, non-split, e < 16 only)")
    n = len(off1) - 1
    out = np.zeros(n * params.max_num_best_mappings, dtype=PAIRS_RECORD if params.output_format == 5 else PE_RECORD)
    tr = np.zeros(n, dtype=TRACE) if trace else None
    seq1 = np.ascontiguousarray(seq1, dtype=np.uint8); seq2 = np.ascontiguousarray(seq2, dtype=np.uint8)
    off1 = np.ascontiguousarray(off1, dtype=np.uint32); off2 = np.ascontiguousarray(off2, dtype=np.uint32)
    got = L.orc_map_pairs_mt(m, n, seq1.ctypes.data, off1.ctypes.data, seq2.ctypes.data, off2.ctypes.data,
                             first_read_id, out.ctypes.data, len(out), n_threads, tr.ctypes.data if trace else None)
    L.orc_mapper_free(m)
    return out[:got], tr


def postprocess(params, recs):
    recs = np.ascontiguousarray(recs.copy())
    n = lib().orc_postprocess(C.byref(params), recs.ctypes.data, len(recs))
    return recs[:n]


def format_bed(ref, recs):
    recs = np.ascontiguousarray(recs)
    n = lib().orc_format_bed(ref.h, recs.ctypes.data, len(recs), None, 0)
    buf = C.create_string_buffer(n + 1)
    lib().orc_format_bed(ref.h, recs.ctypes.data, len(recs), buf, n)
    return buf.raw[:n]


def format_tagalign(ref, recs):
    recs = np.ascontiguousarray(recs)
    n = lib().orc_format_tagalign(ref.h, recs.ctypes.data, len(recs), None, 0)
    buf = C.create_string_buffer(n + 1)
    lib().orc_format_tagalign(ref.h, recs.ctypes.data, len(recs), buf, n)
    return buf.raw[:n]


def run_files(params, index_path, ref_path, r1, r2, out, n_threads=1):
    secs = C.c_double()
    n = C.c_uint64()
    rc = lib().orc_run_files(C.byref(params), index_path.encode(), ref_path.encode(), r1.encode(), r2.encode(),
                             out.encode(), n_threads, C.byref(secs), C.byref(n))
    if rc != 0:
        raise RuntimeError("orc_run_files failed: %d" % rc)
    return secs.value, n.value


def map_reads_se(params, index, ref, seq, off, first_read_id=0, n_threads=1):
    """Single-end: one batch of reads -> records (alignment lengths unused)."""
    L = lib()
    m = L.orc_mapper_create(C.byref(params), index.h, ref.h)
    if not m:
        raise ValueError("unsupported parameters for the oracle")
    n = len(off) - 1
    out = np.zeros(n * params.max_num_best_mappings, dtype=PE_RECORD)
    seq = np.ascontiguousarray(seq, dtype=np.uint8); off = np.ascontiguousarray(off, dtype=np.uint32)
    got = L.orc_map_reads_se(m, n, seq.ctypes.data, off.ctypes.data, first_read_id, out.ctypes.data, len(out), n_threads)
    L.orc_mapper_free(m)
    return out[:got]


def map_reads_se_bc(params, index, ref, seq, off, barcodes, quals, bc_len, whitelist=None, first_read_id=0, n_threads=1):
    L = lib()
    m = L.orc_mapper_create(C.byref(params), index.h, ref.h)
    if whitelist is not None:
        L.orc_mapper_set_barcodes(m, whitelist.h, 1, 0.9, 0)
    n = len(off) - 1
    out = np.zeros(n * params.max_num_best_mappings, dtype=PE_RECORD)
    obc = np.zeros(len(out), dtype=np.uint64)
    st = np.zeros(2, dtype=np.uint64)
    seq = np.ascontiguousarray(seq, dtype=np.uint8); off = np.ascontiguousarray(off, dtype=np.uint32)
    barcodes = np.ascontiguousarray(barcodes, dtype=np.uint8); quals = np.ascontiguousarray(quals, dtype=np.uint8)
    got = L.orc_map_reads_se_bc(m, n, seq.ctypes.data, off.ctypes.data, barcodes.ctypes.data, quals.ctypes.data, bc_len, first_read_id,
                                out.ctypes.data, obc.ctypes.data, len(out), n_threads, st.ctypes.data)
    L.orc_mapper_free(m)
    return out[:got], obc[:got], st


def postprocess_bc(params, recs, bcs):
    recs = np.ascontiguousarray(recs.copy()); bcs = np.ascontiguousarray(bcs.copy(), dtype=np.uint64)
    n = lib().orc_postprocess_bc(C.byref(params), recs.ctypes.data, bcs.ctypes.data, len(recs))
    return recs[:n], bcs[:n]


def format_bed_bc(ref, recs, bcs, bc_len):
    recs = np.ascontiguousarray(recs); bcs = np.ascontiguousarray(bcs, dtype=np.uint64)
    n = lib().orc_format_bed_bc(ref.h, recs.ctypes.data, bcs.ctypes.data, len(recs), bc_len, None, 0)
    buf = C.create_string_buffer(n + 1)
    lib().orc_format_bed_bc(ref.h, recs.ctypes.data, bcs.ctypes.data, len(recs), bc_len, buf, n)
    return buf.raw[:n]


def postprocess_se(params, recs):
    recs = np.ascontiguousarray(recs.copy())
    n = lib().orc_postprocess_se(C.byref(params), recs.ctypes.data, len(recs))
    return recs[:n]


def run_files_se(params, index_path, ref_path, r1, out, n_threads=1):
    rc = lib().orc_run_files_se(C.byref(params), index_path.encode(), ref_path.encode(), r1.encode(), out.encode(), n_threads)
    if rc != 0:
        raise RuntimeError("orc_run_files_se failed: %d" % rc)


SAM_RECORD = np.dtype([("read_id", "<u4"), ("rid", "<u4"), ("pos", "<u4", 2), ("end", "<u4", 2), ("strand", "u1", 2), ("mapq", "u1"), ("is_unique", "u1"),
                       ("secondary", "u1"), ("n_cigar", "u1", 2), ("overflow", "u1"), ("cigar", "<u4", (2, 24))], align=True)


def map_sam_cores(params, index, ref, seq1, off1, seq2=None, off2=None, first_read_id=0):
    """SAM cores (spans, strands, MAPQ, CIGARs) per reported pair / read, in the layout of cmx_sam_record."""
    L = lib()
    m = L.orc_mapper_create(C.byref(params), index.h, ref.h)
    n = len(off1) - 1
    out = np.zeros(n * params.max_num_best_mappings, dtype=SAM_RECORD)
    seq1 = np.ascontiguousarray(seq1, dtype=np.uint8); off1 = np.ascontiguousarray(off1, dtype=np.uint32)
    if seq2 is not None:
        seq2 = np.ascontiguousarray(seq2, dtype=np.uint8); off2 = np.ascontiguousarray(off2, dtype=np.uint32)
    got = L.orc_map_sam_cores(m, n, seq1.ctypes.data, off1.ctypes.data, seq2.ctypes.data if seq2 is not None else None,
                              off2.ctypes.data if seq2 is not None else None, first_read_id, out.ctypes.data, len(out))
    L.orc_mapper_free(m)
    return out[:got]


def run_files_sam(params, index_path, ref_path, r1, r2, out):
    rc = lib().orc_run_files_sam(C.byref(params), index_path.encode(), ref_path.encode(), r1.encode(), (r2 or "").encode(), out.encode())
    if rc != 0:
        raise RuntimeError("orc_run_files_sam failed: %d" % rc)


def run_files_paf(params, index_path, ref_path, r1, r2, out):
    rc = lib().orc_run_files_paf(C.byref(params), index_path.encode(), ref_path.encode(), r1.encode(), (r2 or "").encode(), out.encode())
    if rc != 0:
        raise RuntimeError("orc_run_files_paf failed: %d" % rc)


class Whitelist:
    """Barcode whitelist + abundances (chromap.cc:388-548)."""

    def __init__(self, path, bc_len):
        self.h = lib().orc_whitelist_load(path.encode(), bc_len)
        if not self.h:
            raise IOError("cannot load whitelist")
        self.bc_len = bc_len

    def sample(self, barcodes, max_sample=20000000, batch=500000):
        barcodes = np.ascontiguousarray(barcodes, dtype=np.uint8)
        lib().orc_whitelist_sample(self.h, barcodes.ctypes.data, len(barcodes) // self.bc_len, self.bc_len, max_sample, batch)

    def arrays(self):
        k, c, ns = C.c_void_p(), C.c_void_p(), C.c_uint64()
        n = lib().orc_whitelist_arrays(self.h, C.byref(k), C.byref(c), C.byref(ns))
        keys = np.ctypeslib.as_array(C.cast(k, C.POINTER(C.c_uint64)), (n,)).copy()
        counts = np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_uint32)), (n,)).copy()
        return keys, counts, ns.value


def map_pairs_bc(params, index, ref, seq1, off1, seq2, off2, barcodes, quals, bc_len, whitelist=None, first_read_id=0, n_threads=1):
    L = lib()
    m = L.orc_mapper_create(C.byref(params), index.h, ref.h)
    if whitelist is not None:
        L.orc_mapper_set_barcodes(m, whitelist.h, 1, 0.9, 0)
    n = len(off1) - 1
    out = np.zeros(n * params.max_num_best_mappings, dtype=PE_RECORD)
    obc = np.zeros(len(out), dtype=np.uint64)
    st = np.zeros(2, dtype=np.uint64)
    seq1 = np.ascontiguousarray(seq1, dtype=np.uint8); seq2 = np.ascontiguousarray(seq2, dtype=np.uint8)
    off1 = np.ascontiguousarray(off1, dtype=np.uint32); off2 = np.ascontiguousarray(off2, dtype=np.uint32)
    barcodes = np.ascontiguousarray(barcodes, dtype=np.uint8); quals = np.ascontiguousarray(quals, dtype=np.uint8)
    got = L.orc_map_pairs_bc(m, n, seq1.ctypes.data, off1.ctypes.data, seq2.ctypes.data, off2.ctypes.data, barcodes.ctypes.data, quals.ctypes.data,
                             bc_len, first_read_id, out.ctypes.data, obc.ctypes.data, len(out), n_threads, st.ctypes.data)
    L.orc_mapper_free(m)
    return out[:got], obc[:got], st


def run_files_bc(params, index_path, ref_path, r1, r2, barcode_path, whitelist_path, out, n_threads=1):
    st = np.zeros(2, dtype=np.uint64)
    rc = lib().orc_run_files_bc(C.byref(params), index_path.encode(), ref_path.encode(), r1.encode(), r2.encode(), barcode_path.encode(),
                                (whitelist_path or "").encode(), out.encode(), n_threads, st.ctypes.data)
    if rc != 0:
        raise RuntimeError("orc_run_files_bc failed: %d" % rc)
    return st
