"""-bwt (first "next" row): the `.bwt` file of
bwttab2file (/root/reference/src/match/sfx-run.c:173-210).
CPU: the oracle's bwt image against the files of the unmodified reference
(tests/golden/bwt_vectors.npz, made by tests/golden/make_golden_bwt.py).
GPU: the CUDA path (k_bwt through gtb_esa_copy_bwttab) against both."""
import hashlib
import os
import numpy as np
import pytest

import esa_oracle as eo
import synth
from conftest import ROOT
from genometools_b200 import encode_symbols

Z = np.load(os.path.join(ROOT, "tests", "golden", "bwt_vectors.npz"), allow_pickle=False)
CASES = [str(c) for c in Z["__cases__"]]


def case_input(golden, case):
    m = golden.meta(case)
    return golden.symbols(case), m


@pytest.mark.parametrize("case", CASES)
def test_oracle_bwt_matches_reference(golden, case):
    sym, m = case_input(golden, case)
    o = eo.esa(sym, m["numofchars"], m["prefixlength"])
    assert hashlib.md5(eo.file_images(o)["suf"]).hexdigest() == str(Z[case + "/md5_suf"])   # same case
    got = eo.bwt_image(o, sym)
    assert got == bytes(Z[case + "/bwt"])
    assert hashlib.md5(got).hexdigest() == str(Z[case + "/md5_bwt"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_bwt_matches_reference(golden, case):
    from genometools_b200.suffixerator import build_esa
    sym, m = case_input(golden, case)
    enc = encode_symbols(sym, m["numofchars"], m["numofsequences"])
    res = build_esa(enc, m["prefixlength"], want_bwt=True)
    assert res.bwt_bytes() == bytes(Z[case + "/bwt"])


@pytest.mark.gpu
def test_cuda_bwt_parts_and_cli(tmp_path):
    from genometools_b200.suffixerator import build_esa, suffixerator_main
    sym = synth.reads(300, 80, 5, p_n=0.02)
    o = eo.esa(sym, 4, 4)
    ref = eo.bwt_image(o, sym)
    for parts in (1, 3):
        res = build_esa(encode_symbols(sym, 4), 4, parts=parts, want_bwt=True)
        assert res.bwt_bytes() == ref, parts
    fa = tmp_path / "r.fa"
    synth.to_fasta(sym, str(fa), "dna")
    assert suffixerator_main(["-dna", "-suf", "-bwt", "-pl", "4", "-db", str(fa), "-indexname", str(tmp_path / "idx")]) == 0
    assert (tmp_path / "idx.bwt").read_bytes() == ref
