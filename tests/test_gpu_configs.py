"""GPU: the BASELINE.json configurations AT FULL SIZE against the unmodified reference.

tests/golden/config_md5.json holds, for c2..c5 as genometools_b200/synthetic.py generates
them, length + md5 + order-dependent checksum (genometools_b200/mixhash.py) of the
.suf/.lcp/.llv/.bck files `gtref suffixerator -suf -lcp -bck -pl` wrote here
(tests/golden/make_golden_configs.py; the reference run of c4 takes most of an hour).

  * library path (C-ABI): the tables are copied to the host exactly as the files hold
    them (uint64 suftab, uint8 lcptab, llv pairs, uint32 bucket tables padded to 8 bytes)
    and their md5 must equal the reference's -- byte identity at configuration size;
    the checksums computed over the tables in HBM (gtb_esa_hash_results) must equal the
    reference's checksums too (that is the check bench.py repeats at 1/2/4/8 GPUs).
  * drop-in binary: host/_build/gt_b200 on the c2 FASTA, md5 of all five files.

c4 (3.1 Gbp: 28 GB of host tables, 85 GB of HBM) is compared through the checksums only
unless GTB_TEST_C4_MD5=1.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import synth
from conftest import ROOT
from genometools_b200 import _lib, synthetic as sy
from genometools_b200._lib import GtbStats, GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, ptr
from genometools_b200.mixhash import mixhash
from genometools_b200.suffixerator import recommendedprefixlength

pytestmark = pytest.mark.gpu
GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "config_md5.json")
GOLDEN = json.load(open(GOLDEN_PATH)) if os.path.exists(GOLDEN_PATH) else {}
GT_B200 = os.path.join(ROOT, "host", "_build", "gt_b200")


def md5(a):
    h = hashlib.md5()
    mv = memoryview(np.ascontiguousarray(a)).cast("B")
    for o in range(0, len(mv), 1 << 28):
        h.update(mv[o:o + (1 << 28)])
    return h.hexdigest()


def bck_image(lb, csc, dist):
    def pad(t):
        b = t.tobytes()
        return b + b"\0" * (-len(b) % 8)
    return pad(lb) + pad(csc) + pad(dist)


def run_config(key, want_md5, overlapped=False):
    """overlapped: gtb_esa_run_to_host -- the suffix table crosses PCIe while the ties are refined, the
    entries that left too early are patched -- instead of gtb_esa_run + gtb_esa_copy_results"""
    g = GOLDEN[key]
    w = sy.make_workload(g["workload"], g["scale"])
    n = w.totallength
    assert n == g["totallength"]
    if w.is_dna:
        assert mixhash(w.words) == g["input"]["words"] and mixhash(w.ranges) == g["input"]["ranges"], \
            "the generator does not reproduce the input the golden checksums were made from"
    else:
        assert mixhash(w.symbols) == g["input"]["symbols"]
    pl = recommendedprefixlength(w.numofchars, n)
    assert pl == g["prefixlength"]
    lib = _lib.load()
    buf = C.create_string_buffer(512)
    h = lib.gtb_esa_new(0, buf, 512)
    assert h, buf.value.decode()
    try:
        def ck(rc):
            assert rc == 0, lib.gtb_esa_error(h).decode()
        if w.is_dna:
            ck(lib.gtb_esa_set_input_2bit(h, ptr(w.words), w.words.shape[0], n,
                                          ptr(w.ranges) if w.ranges.shape[0] else None, w.ranges.shape[0]))
        else:
            ck(lib.gtb_esa_set_input_bytes(h, ptr(w.symbols), n, w.numofchars))
        host = None
        if overlapped:
            e = n + 1
            a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
            lib.gtb_bck_sizes(w.numofchars, pl, C.byref(a), C.byref(b), C.byref(c))
            kmax = g["files"]["llv"]["bytes"] // 16 + 8
            host = {"suf": np.empty(e, dtype=np.uint64), "lcp": np.empty(e, dtype=np.uint8),
                    "llv": np.empty(2 * kmax, dtype=np.uint64), "lb": np.empty(a.value + 1, dtype=np.uint32),
                    "csc": np.empty(b.value, dtype=np.uint32), "dist": np.empty(c.value, dtype=np.uint32)}
            nllv = C.c_uint64()
            ck(lib.gtb_esa_run_to_host(h, pl, GTB_WANT_SUF | GTB_WANT_LCP | GTB_WANT_BCK, ptr(host["suf"]), ptr(host["lcp"]),
                                       ptr(host["llv"]), kmax, C.byref(nllv), ptr(host["lb"]),
                                       ptr(host["csc"]) if b.value else None, ptr(host["dist"]) if c.value else None))
            host["k"] = int(nllv.value)
        else:
            ck(lib.gtb_esa_run(h, pl, GTB_WANT_SUF | GTB_WANT_LCP | GTB_WANT_BCK))
        st = GtbStats()
        ck(lib.gtb_esa_get_stats(h, C.byref(st)))
        # the tables in HBM against the reference's files
        out3 = (C.c_uint64 * 3)()
        ck(lib.gtb_esa_hash_results(h, 0, out3))
        hb = C.c_uint64()
        ck(lib.gtb_esa_hash_bcktab(h, C.byref(hb)))
        files = g["files"]
        assert out3[0] == files["suf"]["mixhash"], "suftab in HBM differs from the reference's .suf"
        assert out3[1] == files["lcp"]["mixhash"], "lcptab in HBM differs from the reference's .lcp"
        assert out3[2] == files["llv"]["mixhash"], "llv pairs in HBM differ from the reference's .llv"
        assert hb.value == files["bck"]["mixhash"], "bucket table in HBM differs from the reference's .bck"
        assert lib.gtb_esa_num_entries(h) == n + 1
        assert 16 * lib.gtb_esa_num_llv(h) == files["llv"]["bytes"]
        # the .prj lines the sorter is responsible for
        prj = dict(line.split("=", 1) for line in g["prj"].strip().split("\n"))
        assert st.longest == int(prj["longest"])
        assert st.numoflargelcpvalues == int(prj["largelcpvalues"])
        assert st.maxbranchdepth == int(prj["maxbranchdepth"])
        assert "%.2f" % (st.lcptabsum / (n + 1)) == prj["averagelcp"]
        assert st.specialcharacters == int(prj["specialcharacters"])
        if host is not None:
            k = host["k"]
            assert md5(host["suf"]) == files["suf"]["md5"], ".suf (overlapped copy)"
            assert md5(host["lcp"]) == files["lcp"]["md5"], ".lcp (overlapped copy)"
            assert 16 * k == files["llv"]["bytes"] and md5(host["llv"][:2 * k]) == files["llv"]["md5"], ".llv"
            img = bck_image(host["lb"], host["csc"], host["dist"])
            assert hashlib.md5(img).hexdigest() == files["bck"]["md5"], ".bck"
        elif want_md5:
            e = n + 1
            k = int(lib.gtb_esa_num_llv(h))
            suf = np.empty(e, dtype=np.uint64)
            lcp = np.empty(e, dtype=np.uint8)
            llv = np.empty(2 * max(k, 1), dtype=np.uint64)
            a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
            lib.gtb_bck_sizes(w.numofchars, pl, C.byref(a), C.byref(b), C.byref(c))
            lb = np.empty(a.value + 1, dtype=np.uint32)
            csc = np.empty(b.value, dtype=np.uint32)
            dist = np.empty(c.value, dtype=np.uint32)
            ck(lib.gtb_esa_copy_results(h, ptr(suf), ptr(lcp), ptr(llv) if k else None, ptr(lb),
                                        ptr(csc) if b.value else None, ptr(dist) if c.value else None))
            assert suf.nbytes == files["suf"]["bytes"] and md5(suf) == files["suf"]["md5"], ".suf"
            assert md5(lcp) == files["lcp"]["md5"], ".lcp"
            assert md5(llv[:2 * k]) == files["llv"]["md5"], ".llv"
            img = bck_image(lb, csc, dist)
            assert len(img) == files["bck"]["bytes"] and hashlib.md5(img).hexdigest() == files["bck"]["md5"], ".bck"
    finally:
        lib.gtb_esa_delete(h)


def need(key):
    return pytest.mark.skipif(key not in GOLDEN, reason=f"no golden checksums for {key} (make_golden_configs.py)")


@need("c2")
def test_c2_full_size_byte_identical():
    run_config("c2", True)


@need("c5")
def test_c5_full_size_byte_identical():
    run_config("c5", True)


@need("c3")
def test_c3_full_size_byte_identical_overlapped_copy():
    """c3 through gtb_esa_run_to_host (c2 and c5 above: gtb_esa_run + gtb_esa_copy_results)"""
    run_config("c3", True, overlapped=True)


@need("c4@0.002")
def test_overlapped_copy_patches_small_chunks(monkeypatch):
    """64 KB staging chunks: the early copy runs on a 6.2 Mbp sample full of ties, every chunk is patched"""
    monkeypatch.setenv("GTB200_STAGE_KB", "64")
    run_config("c4@0.002", True, overlapped=True)


@need("c4")
def test_c4_full_size_identical_checksums():
    run_config("c4", os.environ.get("GTB_TEST_C4_MD5") == "1")


@need("c2")
@pytest.mark.skipif(not os.path.exists(GT_B200), reason="host/_build/gt_b200 not built")
def test_c2_dropin_binary_files_md5(tmp_path):
    """`gt_b200 suffixerator -dna -suf -lcp -bck -pl` on the 100 Mbp FASTA: all five index files
    have the md5 of the files the unmodified reference wrote"""
    g = GOLDEN["c2"]
    w = sy.make_workload("c2", 1.0)
    fa = str(tmp_path / "c2.fa")
    synth.to_fasta(w.to_symbols(), fa, "dna")
    idx = str(tmp_path / "c2")
    subprocess.check_call([GT_B200, "suffixerator", "-dna", "-suf", "-lcp", "-bck", "-pl", "-indexname", idx,
                           "-db", fa], stdout=subprocess.DEVNULL)
    for ext in ("suf", "lcp", "llv", "bck"):
        h = hashlib.md5()
        with open(f"{idx}.{ext}", "rb") as fh:
            for blk in iter(lambda: fh.read(1 << 26), b""):
                h.update(blk)
        assert h.hexdigest() == g["files"][ext]["md5"], ext
    # the .prj names no file: identical text
    assert open(idx + ".prj").read() == g["prj"]
