"""-dir rev|cpl|rcl (first "next" row): the sequence read in reverse /
complement direction (/root/reference/src/core/encseq.c:6094-6140,
src/match/sfx-mapped4.gen:33-86).
CPU: the oracle on the transformed symbols against the files of the unmodified reference
(tests/golden/readmode_vectors.npz, made by tests/golden/make_golden_readmode.py).
GPU: the CUDA path (gtb_esa_set_readmode: k_readmode_words / k_readmode_bytes) against them."""
import hashlib
import os
import numpy as np
import pytest

import esa_oracle as eo
import synth
from conftest import ROOT
from genometools_b200 import encode_symbols

Z = np.load(os.path.join(ROOT, "tests", "golden", "readmode_vectors.npz"), allow_pickle=False)
CASES = [str(c) for c in Z["__cases__"]]
EXTS = ("suf", "lcp", "llv", "bck", "bwt")


def md5(b):
    return hashlib.md5(b).hexdigest()


def split(case):
    name, mode = case.rsplit("@", 1)
    return name, mode


@pytest.mark.parametrize("case", CASES)
def test_oracle_readmode_matches_reference(golden, case):
    name, mode = split(case)
    m = golden.meta(name)
    sym = eo.apply_readmode(golden.symbols(name), mode)
    pl = int(Z[case + "/prefixlength"])
    o = eo.esa(sym, m["numofchars"], pl)
    im = eo.file_images(o)
    im["bwt"] = eo.bwt_image(o, sym)
    for ext in EXTS:
        assert len(im[ext]) == int(Z[f"{case}/len_{ext}"]) and md5(im[ext]) == str(Z[f"{case}/md5_{ext}"]), ext
    prj = bytes(Z[case + "/prj"]).decode()
    for line in eo.prj_sorter_lines(o):
        assert line in prj.split("\n"), line


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_readmode_matches_reference(golden, case):
    from genometools_b200.suffixerator import build_esa
    name, mode = split(case)
    m = golden.meta(name)
    enc = encode_symbols(golden.symbols(name), m["numofchars"], m["numofsequences"])
    res = build_esa(enc, int(Z[case + "/prefixlength"]), want_bwt=True, readmode=mode)
    got = {"suf": res.suf_bytes(), "lcp": res.lcp_bytes(), "llv": res.llv_bytes(), "bck": res.bck_bytes(),
           "bwt": res.bwt_bytes()}
    for ext in EXTS:
        assert len(got[ext]) == int(Z[f"{case}/len_{ext}"]) and md5(got[ext]) == str(Z[f"{case}/md5_{ext}"]), ext
    assert res.prj_text(enc.specialcharinfo(), enc.numofsequences) == bytes(Z[case + "/prj"]).decode()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 63, 64, 65, 1000, 4097])
@pytest.mark.parametrize("mode", ["rev", "cpl", "rcl"])
def test_cuda_readmode_word_boundaries(n, mode):
    """lengths around the 32-base word size, with specials at both ends, plus -parts"""
    from genometools_b200.suffixerator import build_esa
    rng = np.random.default_rng(n * 7 + len(mode))
    sym = rng.integers(0, 4, n).astype(np.uint8)
    if n >= 3:
        sym[0] = 254
        sym[n - 1] = 255 if n > 40 else sym[n - 1]
        sym[rng.integers(0, n, max(1, n // 50))] = 254
    t = eo.apply_readmode(sym, mode)
    pl = 2 if n >= 1000 else 1
    o = eo.esa(t, 4, pl)
    im = eo.file_images(o)
    for parts in (1, 2):
        res = build_esa(encode_symbols(sym, 4), pl, parts=parts, want_bwt=True, readmode=mode)
        assert res.suf_bytes() == im["suf"] and res.lcp_bytes() == im["lcp"] and res.bck_bytes() == im["bck"], parts
        assert res.bwt_bytes() == eo.bwt_image(o, t), parts
        assert res.longest == o["longest"]


@pytest.mark.gpu
def test_cuda_readmode_bytes_path_and_errors():
    from genometools_b200.suffixerator import build_esa
    from genometools_b200._lib import GtbError
    sym = synth.protein(30_000, 4)
    t = eo.apply_readmode(sym, "rev")
    o = eo.esa(t, 20, 2)
    res = build_esa(encode_symbols(sym, 20), 2, readmode="rev", want_bwt=True)
    im = eo.file_images(o)
    assert res.suf_bytes() == im["suf"] and res.lcp_bytes() == im["lcp"] and res.bck_bytes() == im["bck"]
    assert res.bwt_bytes() == eo.bwt_image(o, t)
    with pytest.raises(GtbError, match="only can be used for DNA"):
        build_esa(encode_symbols(sym, 20), 2, readmode="rcl")
    with pytest.raises(GtbError, match="unknown readmode"):
        build_esa(encode_symbols(sym, 20), 2, readmode="sideways")
