"""CPU: the order-dependent checksum (genometools_b200/mixhash.py) and the golden table
tests/golden/config_md5.json (outputs of the unmodified reference on the BASELINE.json
configurations).  The bounded-sample entries are re-derived here with the pinned oracle:
the generator, the reference run, the checksum and the oracle must all agree."""
import hashlib
import json
import os

import numpy as np
import pytest

import esa_oracle as eo
from conftest import ROOT
from genometools_b200 import synthetic as sy
from genometools_b200.mixhash import mixhash, mixhash_file, FILE_DTYPES

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "config_md5.json")))


def test_mixhash_composes_across_shards_and_depends_on_order():
    rng = np.random.default_rng(3)
    v = rng.integers(0, 2 ** 32, size=100_000, dtype=np.uint64)
    whole = mixhash(v)
    cuts = [0, 1, 777, 50_000, 99_999, 100_000]
    assert sum(mixhash(v[a:b], a) for a, b in zip(cuts, cuts[1:])) % 2 ** 64 == whole
    assert mixhash(v, chunk=4096) == whole
    w = v.copy(); w[[10, 11]] = w[[11, 10]]
    assert mixhash(w) != whole                     # same multiset, different order
    assert mixhash(v.astype(np.uint32)) == whole   # the entry width does not matter, only index and value
    assert mixhash(np.zeros(0, dtype=np.uint8)) == 0


def test_mixhash_known_answers():
    # fixed vectors: the device kernel (k_mixhash) is tested against the same function on the GPU box
    assert mixhash(np.array([0], dtype=np.uint8)) == 0xccd8a7449c0ac4ba
    assert mixhash(np.arange(10, dtype=np.uint64)) == 0x34a89d716d7786e9
    assert mixhash(np.arange(10, dtype=np.uint8), 5) == 0x371eb91ec12759a


def test_mixhash_file(tmp_path):
    rng = np.random.default_rng(4)
    v = rng.integers(0, 2 ** 63, size=70_001, dtype=np.uint64)
    p = tmp_path / "t.bin"
    v.astype("<u8").tofile(p)
    assert mixhash_file(str(p), "<u8", chunk_bytes=8 * 1000) == mixhash(v)


def test_golden_table_is_complete():
    for key in ("c2", "c2@0.01", "c3@0.001", "c4@0.0005", "c5@0.002"):
        assert key in GOLDEN, key
    for key, g in GOLDEN.items():
        n = g["totallength"]
        assert g["files"]["suf"]["bytes"] == 8 * (n + 1) and g["files"]["lcp"]["bytes"] == n + 1, key
        prj = dict(line.split("=", 1) for line in g["prj"].strip().split("\n"))
        assert int(prj["totallength"]) == n and int(prj["largelcpvalues"]) * 16 == g["files"]["llv"]["bytes"], key
        assert int(prj["prefixlength"]) == g["prefixlength"], key


@pytest.mark.parametrize("key", ["c2@0.01", "c3@0.001", "c4@0.0005", "c5@0.002"])
def test_oracle_reproduces_golden_config_samples(key):
    g = GOLDEN[key]
    w = sy.make_workload(g["workload"], g["scale"])
    if w.is_dna:
        assert mixhash(w.words) == g["input"]["words"] and mixhash(w.ranges) == g["input"]["ranges"]
    else:
        assert mixhash(w.symbols) == g["input"]["symbols"]
    o = eo.esa(w.to_symbols(), w.numofchars, g["prefixlength"])
    im = eo.file_images(o)
    for ext, dt in FILE_DTYPES.items():
        assert len(im[ext]) == g["files"][ext]["bytes"], ext
        assert hashlib.md5(im[ext]).hexdigest() == g["files"][ext]["md5"], ext
        assert mixhash(np.frombuffer(im[ext], dtype=dt)) == g["files"][ext]["mixhash"], ext
