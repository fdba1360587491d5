"""GPU: the look-back of the onesweep passes with the formally ordered memory operations.

The library polls the per-tile status words with weak L1-bypassing loads first (they pipeline) and
falls back to ld.relaxed.gpu; GTB_RS_STRONG compiles every status access as st.relaxed.gpu /
ld.relaxed.gpu (gtb_radix.cuh).  This test builds that variant of the library next to the product
one and checks that both produce the reference's tables."""
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from genometools_b200 import synthetic as sy
from genometools_b200._lib import ptr

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "config_md5.json")))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_strong_lookback_library_matches_reference(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    so = str(tmp_path / "libgtb200_strong.so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
                           "-Xcompiler", "-fPIC", "-DGTB_RS_STRONG", "-o", so,
                           os.path.join(ROOT, "genometools_b200", "csrc", "gtb_esa.cu"),
                           os.path.join(ROOT, "genometools_b200", "csrc", "gtb_widen.cpp")])
    lib = C.CDLL(so)
    lib.gtb_esa_new.restype = C.c_void_p
    lib.gtb_esa_error.restype = C.c_char_p
    for f in ("gtb_esa_delete", "gtb_esa_set_input_2bit", "gtb_esa_run", "gtb_esa_hash_results", "gtb_esa_hash_bcktab",
              "gtb_esa_error"):
        getattr(lib, f).argtypes = None
    key = "c4@0.002"
    g = GOLDEN[key]
    w = sy.make_workload(g["workload"], g["scale"])
    buf = C.create_string_buffer(512)
    h = C.c_void_p(lib.gtb_esa_new(0, buf, 512))
    assert h.value, buf.value.decode()
    try:
        rc = lib.gtb_esa_set_input_2bit(h, C.c_void_p(ptr(w.words)), C.c_uint64(w.words.shape[0]), C.c_uint64(w.totallength),
                                        C.c_void_p(ptr(w.ranges)), C.c_uint64(w.ranges.shape[0]))
        assert rc == 0, lib.gtb_esa_error(h).decode()
        assert lib.gtb_esa_run(h, C.c_uint(g["prefixlength"]), C.c_uint(7)) == 0, lib.gtb_esa_error(h).decode()
        out3 = (C.c_uint64 * 3)()
        hb = C.c_uint64()
        assert lib.gtb_esa_hash_results(h, C.c_uint64(0), out3) == 0
        assert lib.gtb_esa_hash_bcktab(h, C.byref(hb)) == 0
        got = {"suf": out3[0], "lcp": out3[1], "llv": out3[2], "bck": hb.value}
        for ext in ("suf", "lcp", "llv", "bck"):
            assert got[ext] == g["files"][ext]["mixhash"], ext
    finally:
        lib.gtb_esa_delete(h)
