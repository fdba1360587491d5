"""CPU: the oracle restatement (oracle/esa_oracle.c) against the outputs of the
unmodified reference (tests/golden/reference_vectors.npz, made by
tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import numpy as np
import pytest

import esa_oracle as eo
from conftest import golden_cases


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference(golden, case):
    m = golden.meta(case)
    o = eo.esa(golden.symbols(case), m["numofchars"], m["prefixlength"])
    im = eo.file_images(o)
    for ext in ("suf", "lcp", "llv", "bck"):
        assert len(im[ext]) == int(golden.get(case, "len_" + ext)), ext
        assert hashlib.md5(im[ext]).hexdigest() == str(golden.get(case, "md5_" + ext)), ext
    prj, text = golden.prj(case)
    for line in eo.prj_sorter_lines(o):
        assert line + "\n" in text, line
    assert int(prj["specialcharacters"]) == o["specialcharacters"]


def test_golden_covers_the_reference_fixture_list(golden):
    # the 25 fixtures of testsuite/gt_suffixerator_include.rb:119-143 + protein + multi-file + synthetic
    names = {c.split("/")[1] for c in golden.cases if c.startswith("file/")}
    assert len(names) == 27
    assert any(int(golden.get(c, "len_llv")) > 0 for c in golden.cases)      # .llv exercised
    assert any(golden.meta(c)["alphabet"] == "protein" for c in golden.cases)


def test_oracle_edge_cases():
    # empty text, only specials, single symbol
    o = eo.esa(np.zeros(0, np.uint8), 4, 1)
    assert o["suf"].tolist() == [0] and o["lcp"].tolist() == [0]
    o = eo.esa(np.array([254, 255, 254], np.uint8), 4, 1)
    assert o["suf"].tolist() == [0, 1, 2, 3] and o["leftborder"].tolist() == [0, 0, 0, 0, 0]
    o = eo.esa(np.array([3], np.uint8), 4, 2)
    assert o["suf"].tolist() == [0, 1] and o["longest"] == 0
    assert o["leftborder"][-1] == 1 and o["countspecialcodes"][3] == 1
