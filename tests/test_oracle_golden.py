"""The CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED reference binary
(tests/golden/make_golden.sh).  CPU only."""
import gzip
import hashlib
import os

import pytest

from oracle import oracle_py as orc
from tests.util import load_pairs


def _md5s(d):
    out = {}
    for line in open(os.path.join(d, "md5.txt")):
        h, name = line.split()
        out[name] = h
    return out


CASES = {
    "chip": dict(preset="chip"),
    "atac": dict(preset="atac"),
    "default": dict(preset=""),
    "q0dedup": dict(preset="", remove_pcr_duplicates=1, mapq_threshold=0),
    "e5": dict(preset="", error_threshold=5, mapq_threshold=10, tn5_shift=1, remove_pcr_duplicates=1),
    "e12l300": dict(preset="", error_threshold=12, max_insert_size=300, mapq_threshold=0),
    "n3q0": dict(preset="", max_num_best_mappings=3, mapq_threshold=0),
}


@pytest.fixture(scope="module")
def synth_index(golden_dir, tmp_path_factory):
    d = os.path.join(golden_dir, "synth_small")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)  # our builder; lookups must equal the reference's khash index
    p = str(tmp_path_factory.mktemp("idx") / "synth.index")
    assert idx.save(p) == 0
    return d, p


@pytest.mark.parametrize("case", sorted(CASES))
def test_synth_small_matches_reference_binary(case, synth_index, tmp_path):
    d, index_path = synth_index
    kw = dict(CASES[case])
    params = orc.make_params(kw.pop("preset"), **kw)
    out = str(tmp_path / "o.bed")
    orc.run_files(params, index_path, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"),
                  os.path.join(d, "read2.fq.gz"), out, n_threads=1)
    got = open(out, "rb").read()
    want = gzip.open(os.path.join(d, case + ".bed.gz")).read()
    assert hashlib.md5(want).hexdigest() == _md5s(d)[case + ".bed"]
    assert got == want


@pytest.mark.parametrize("case,preset", [("default", ""), ("chip", "chip"), ("atac", "atac")])
def test_reference_quickstart_data(case, preset, golden_dir, tmp_path):
    """SURVEY.md §4 golden md5s on /root/reference/test data, with the reference-built index file."""
    d = os.path.join(golden_dir, "ref_test")
    out = str(tmp_path / "o.bed")
    orc.run_files(orc.make_params(preset), os.path.join(d, "ref.index"), os.path.join(d, "ref.fa.gz"),
                  os.path.join(d, "read1.fq"), os.path.join(d, "read2.fq"), out)
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == _md5s(d)[case + ".bed"]
    want = {"default": "e311f0a0848edca7196d839f47b3e007", "chip": "e311f0a0848edca7196d839f47b3e007",
            "atac": "63b977e6e8af35be7861f35b6163f714"}[case]
    assert _md5s(d)[case + ".bed"] == want


def test_threads_do_not_change_output(synth_index, tmp_path):
    d, index_path = synth_index
    params = orc.make_params("", remove_pcr_duplicates=1, mapq_threshold=0)
    outs = []
    for t in (1, 4):
        out = str(tmp_path / ("o%d.bed" % t))
        orc.run_files(params, index_path, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"),
                      os.path.join(d, "read2.fq.gz"), out, n_threads=t)
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]


def test_index_builder_equals_reference_index(golden_dir):
    """Our index builder answers every lookup like the khash table the reference binary wrote."""
    import numpy as np
    d = os.path.join(golden_dir, "ref_test")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    theirs = orc.Index(os.path.join(d, "ref.index"))
    ours = orc.Index(ref=ref, k=theirs.k, w=theirs.w)
    a, b = theirs.arrays(), ours.arrays()
    assert a["n_buckets"] == b["n_buckets"] == 32768
    assert np.array_equal(a["occ"], b["occ"])
    import ctypes as C
    h, _ = orc.minimizers(ref.seq(0), theirs.k, theirs.w)
    assert len(h) == 25079
    L = orc.lib()
    for x in h[::7]:
        k1, v1, k2, v2 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        assert L.orc_index_lookup(theirs.h, int(x), C.byref(k1), C.byref(v1)) == 1
        assert L.orc_index_lookup(ours.h, int(x), C.byref(k2), C.byref(v2)) == 1
        assert (k1.value, v1.value) == (k2.value, v2.value)


HIC_CASES = {
    "hic": dict(),
    "hic_q0": dict(mapq_threshold=0),
    "hic_e6dedup": dict(mapq_threshold=0, error_threshold=6, remove_pcr_duplicates=1),
}


@pytest.mark.parametrize("case", sorted(HIC_CASES))
def test_hic_split_alignment_pairs_match_reference_binary(case, golden_dir, tmp_path):
    d = os.path.join(golden_dir, "synth_hic")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    ip = str(tmp_path / "hic.index")
    assert idx.save(ip) == 0
    out = str(tmp_path / "o.pairs")
    orc.run_files(orc.make_params("hic", **HIC_CASES[case]), ip, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"),
                  os.path.join(d, "read2.fq.gz"), out)
    want = gzip.open(os.path.join(d, case + ".pairs.gz")).read()
    assert hashlib.md5(want).hexdigest() == _md5s(d)[case + ".pairs"]
    assert open(out, "rb").read() == want


def test_hic_reference_quickstart(golden_dir, tmp_path):
    d = os.path.join(golden_dir, "ref_test")
    out = str(tmp_path / "o.pairs")
    orc.run_files(orc.make_params("hic"), os.path.join(d, "ref.index"), os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq"),
                  os.path.join(d, "read2.fq"), out)
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == "fc844a251ebdcec0f641b59fef804d0f"


@pytest.mark.parametrize("case,use_wl", [("sc_whitelist", True), ("sc_nowhitelist", False)])
def test_scatac_barcodes_match_reference_binary(case, use_wl, golden_dir, tmp_path):
    """--preset atac with -b (and --barcode-whitelist): barcode correction, cell-level dedup, barcode BED column."""
    d = os.path.join(golden_dir, "synth_sc")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    ip = str(tmp_path / "sc.index")
    assert idx.save(ip) == 0
    out = str(tmp_path / "o.bed")
    st = orc.run_files_bc(orc.make_params("atac"), ip, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"), os.path.join(d, "read2.fq.gz"),
                          os.path.join(d, "barcode.fq.gz"), os.path.join(d, "whitelist.txt") if use_wl else "", out)
    want = gzip.open(os.path.join(d, case + ".bed.gz")).read()
    assert hashlib.md5(want).hexdigest() == _md5s(d)[case + ".bed"]
    assert open(out, "rb").read() == want
    if use_wl:
        s = open(os.path.join(d, "sc_stats.txt")).read()
        assert "whitelist: %d." % st[0] in s and "corrected barcodes: %d." % st[1] in s


SE_CASES = {
    "se_default": dict(preset=""),
    "se_chip": dict(preset="chip"),
    "se_q0dedup_tn5": dict(preset="", mapq_threshold=0, remove_pcr_duplicates=1, tn5_shift=1),
    "se_n3q0": dict(preset="", max_num_best_mappings=3, mapq_threshold=0),
    "se_lowmem_q0": dict(preset="", low_memory_mode=1, mapq_threshold=0, remove_pcr_duplicates=1, tn5_shift=1),
}


@pytest.mark.parametrize("case", sorted(SE_CASES))
def test_single_end_oracle_reproduces_reference_bed(golden_dir, tmp_path, case):
    """chromap -1 read1.fq (single-end, MappingWithoutBarcode): oracle output == the reference binary's BED."""
    d = os.path.join(golden_dir, "synth_small")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    ip = str(tmp_path / "ref.index")
    idx.save(ip)
    kw = dict(SE_CASES[case])
    p = orc.make_params(kw.pop("preset"), **kw)
    out = str(tmp_path / "out.bed")
    orc.run_files_se(p, ip, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"), out, 2)
    assert open(out, "rb").read() == gzip.open(os.path.join(d, case + ".bed.gz")).read()


def test_tagalign_text_equals_reference(golden_dir):
    """--preset chip --TagAlign: same records as BED, PairedTagAlign text (mapping_writer.cc:84-110)."""
    d = os.path.join(golden_dir, "synth_small")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    s1, o1, s2, o2 = load_pairs(d)
    p = orc.make_params("chip")
    recs, _ = orc.map_pairs(p, idx, ref, s1, o1, s2, o2)
    assert orc.format_tagalign(ref, orc.postprocess(p, recs)) == gzip.open(os.path.join(d, "chip.tagalign.gz")).read()


SE_SC_CASES = {
    "se_sc_whitelist": (dict(preset="atac"), True),
    "se_sc_nowhitelist": (dict(preset="atac"), False),
    "se_sc_inmem": (dict(preset="", mapq_threshold=0, remove_pcr_duplicates=1, tn5_shift=1), True),
}


def _barcodes(path):
    lines = gzip.open(path).read().split(b"\n")
    import numpy as np
    return np.frombuffer(b"".join(lines[1::4]), dtype=np.uint8), np.frombuffer(b"".join(lines[3::4]), dtype=np.uint8), len(lines[1])


@pytest.mark.parametrize("case", sorted(SE_SC_CASES))
def test_single_end_barcoded_oracle_reproduces_reference_bed(golden_dir, case):
    """chromap -1 r1 -b barcodes [--barcode-whitelist]: MappingWithBarcode records, duplicates = same (barcode, start)."""
    d = os.path.join(golden_dir, "synth_sc")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    s1, o1, _, _ = load_pairs(d)
    bcs, quals, bc_len = _barcodes(os.path.join(d, "barcode.fq.gz"))
    kw, use_wl = SE_SC_CASES[case]
    kw = dict(kw)
    p = orc.make_params(kw.pop("preset"), single_end=1, **kw)
    wl = None
    if use_wl:
        wl = orc.Whitelist(os.path.join(d, "whitelist.txt"), bc_len)
        wl.sample(bcs)
    recs, obc, _ = orc.map_reads_se_bc(p, idx, ref, s1, o1, bcs, quals, bc_len, whitelist=wl, n_threads=2)
    r2, b2 = orc.postprocess_bc(p, recs, obc)
    assert orc.format_bed_bc(ref, r2, b2, bc_len) == gzip.open(os.path.join(d, case + ".bed.gz")).read()


SAM_CASES = {
    "pe_chip": (dict(preset="chip"), True),
    "pe_q0d": (dict(preset="", mapq_threshold=0, remove_pcr_duplicates=1), True),
    "se_n3": (dict(preset="", max_num_best_mappings=3, mapq_threshold=0), False),
}


@pytest.mark.parametrize("case", sorted(SAM_CASES))
def test_sam_oracle_reproduces_reference(golden_dir, tmp_path, case):
    """chromap --SAM (ksw_semi_global3 CIGARs, NM / MD, flags, mate fields): groundwork for the round-2 CUDA path; the
    oracle's text equals the reference binary's, header included."""
    d = os.path.join(golden_dir, "synth_small")
    ref = orc.Reference(os.path.join(d, "ref.fa.gz"))
    idx = orc.Index(ref=ref, k=17, w=7)
    ip = str(tmp_path / "ref.index")
    idx.save(ip)
    kw, paired = SAM_CASES[case]
    kw = dict(kw)
    p = orc.make_params(kw.pop("preset"), **kw)
    out = str(tmp_path / "out.sam")
    orc.run_files_sam(p, ip, os.path.join(d, "ref.fa.gz"), os.path.join(d, "read1.fq.gz"), os.path.join(d, "read2.fq.gz") if paired else None, out)
    assert open(out, "rb").read() == gzip.open(os.path.join(d, case + ".sam.gz")).read()
