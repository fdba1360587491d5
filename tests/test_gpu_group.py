"""GPU: one job sharded over several code ranges inside one process (gtb_group, include/gtb200.h).

On a single GPU every range lives on device 0 -- "peer" memory is then local memory, but the code
path is the one that puts one range on each GPU of a box (gt_b200 -j N): coarse count gather and
cut, position-sharded text scan whose partition pass stores into the owners' receive buffers,
first-level sort per range, lock-step prefix doubling with rank lookups read from the owner's rank
map in place, seams, result gather into one table.  With two or more GPUs visible the same cases
also run with one range per device (real peer access over NVLink).

Compared byte for byte with the outputs of the unmodified reference (golden vectors, the
configuration samples of tests/golden/config_md5.json) and with the pinned oracle."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import esa_oracle as eo
import synth
from conftest import ROOT, golden_cases
from genometools_b200 import _lib, encode_symbols, synthetic as sy
from genometools_b200.mixhash import mixhash
from genometools_b200.suffixerator import build_esa, recommendedprefixlength

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "config_md5.json")))


def ndev():
    return _lib.load().gtb_device_count()


def device_lists(nranges):
    out = [[0] * nranges]
    if ndev() >= 2:
        out.append([i % ndev() for i in range(nranges)])
    return out


def images(res):
    return {"suf": res.suf_bytes(), "lcp": res.lcp_bytes(), "llv": res.llv_bytes(), "bck": res.bck_bytes()}


def check_hashes(res, im):
    for ext, dt in (("suf", "<u8"), ("lcp", "u1"), ("llv", "<u8"), ("bck", "<u4")):
        assert res.device_hashes[ext] == mixhash(np.frombuffer(im[ext], dtype=dt)), f"checksum in HBM of .{ext}"


@pytest.mark.parametrize("key", ["c2@0.01", "c3@0.001", "c4@0.0005", "c5@0.002"])
@pytest.mark.parametrize("nranges", [2, 3, 8])
def test_group_matches_reference_on_config_samples(key, nranges):
    g = GOLDEN[key]
    w = sy.make_workload(g["workload"], g["scale"])
    enc = encode_symbols(w.to_symbols(), w.numofchars, w.numofsequences)
    for devs in device_lists(nranges):
        res = build_esa(enc, g["prefixlength"], devices=devs)
        im = images(res)
        for ext in ("bck", "suf", "lcp", "llv"):
            assert len(im[ext]) == g["files"][ext]["bytes"], (ext, devs)
            assert hashlib.md5(im[ext]).hexdigest() == g["files"][ext]["md5"], (ext, devs)
            assert res.device_hashes[ext] == g["files"][ext]["mixhash"], (ext, devs)
        prj = dict(line.split("=", 1) for line in g["prj"].strip().split("\n"))
        assert res.longest == int(prj["longest"]) and res.maxbranchdepth == int(prj["maxbranchdepth"])
        assert res.numoflargelcpvalues == int(prj["largelcpvalues"])
        assert "%.2f" % (res.lcptabsum / (w.totallength + 1)) == prj["averagelcp"]


@pytest.mark.parametrize("case", [c for c in golden_cases(small_only=True) if "/auto" in c or c.startswith("synth/")])
def test_group_matches_reference_golden_vectors(golden, case):
    m = golden.meta(case)
    sym = golden.symbols(case)
    enc = encode_symbols(sym, m["numofchars"], m["numofsequences"])
    res = build_esa(enc, m["prefixlength"], devices=[0, 0, 0])
    got = images(res)
    ref = {"suf": golden.get(case, "suf").astype("<u8").tobytes(), "lcp": bytes(golden.get(case, "lcp")),
           "llv": golden.get(case, "llv").astype("<u8").tobytes(), "bck": bytes(golden.get(case, "bck"))}
    for ext in ("bck", "suf", "lcp", "llv"):
        assert got[ext] == ref[ext], ext
    check_hashes(res, got)


@pytest.mark.parametrize("what", ["polyA", "tiny", "one_symbol", "lowcomplex", "reads_dup", "deep_repeats"])
def test_group_edge_cases_against_oracle(what):
    """fewer parts than ranges (everything in one coarse bucket), tiny texts, ties that need many
    doubling rounds across range borders"""
    if what == "polyA":
        sym, K, pl = np.zeros(5000, dtype=np.uint8), 4, 3
    elif what == "tiny":
        sym, K, pl = np.array([2, 1, 254, 0, 3, 3, 255, 1], dtype=np.uint8), 4, 1
    elif what == "one_symbol":
        sym, K, pl = np.array([1], dtype=np.uint8), 4, 1
    elif what == "lowcomplex":
        sym, K, pl = synth.low_complexity_dna(40_000, 5), 4, 3
    elif what == "reads_dup":
        r = synth.reads(500, 80, 2, 0.0)
        sym, K, pl = np.concatenate([r, [255], r, [255], r]).astype(np.uint8), 4, 5
    else:
        sym, K, pl = synth.repeats_dna(300_000, 3, unit=20_000, copies=8, exact_len=40_000, exact_copies=4), 4, 6
    enc = encode_symbols(sym, K)
    o = eo.esa(sym, K, pl)
    im = eo.file_images(o)
    for nranges in (2, 4, 7):
        for devs in device_lists(nranges):
            res = build_esa(enc, pl, devices=devs)
            got = images(res)
            for ext in ("bck", "suf", "lcp", "llv"):
                assert got[ext] == im[ext], (what, nranges, ext)
            assert res.longest == o["longest"] and res.maxbranchdepth == o["maxbranchdepth"]
            assert res.numoflargelcpvalues == o["numoflargelcpvalues"] and res.lcptabsum == o["lcptabsum"]
            check_hashes(res, got)


def test_group_scan_modes_agree(monkeypatch):
    """GTB200_SHARD_SCAN=filter: every range scans the whole text and keeps the keys of its interval"""
    sym = synth.repeats_dna(200_000, 9, unit=5000, copies=12, exact_len=9000, exact_copies=3)
    enc = encode_symbols(sym, 4)
    a = images(build_esa(enc, 7, devices=[0, 0, 0, 0]))
    monkeypatch.setenv("GTB200_SHARD_SCAN", "filter")
    b = images(build_esa(enc, 7, devices=[0, 0, 0, 0]))
    c = images(build_esa(enc, 7))
    assert a == b == c


def test_group_bwt_readmode_and_protein():
    sym = synth.reads(600, 70, 3, 0.003)
    enc = encode_symbols(sym, 4)
    for mode in ("fwd", "rcl"):
        one = build_esa(enc, 4, want_bwt=True, readmode=mode)
        grp = build_esa(enc, 4, want_bwt=True, readmode=mode, devices=[0, 0, 0])
        assert images(one) == images(grp) and one.bwt_bytes() == grp.bwt_bytes(), mode
    p = synth.protein(30_000, 4)
    encp = encode_symbols(p, 20)
    assert images(build_esa(encp, 2)) == images(build_esa(encp, 2, devices=[0, 0, 0, 0, 0]))


def test_group_repeated_runs_on_one_group():
    """the handles of a group are reused (bench.py steps): a second run on the same group and a
    run on another input give the same tables as fresh groups"""
    lib = _lib.load()
    buf = C.create_string_buffer(512)
    devs = (C.c_int * 3)(0, 0, 0)
    g = lib.gtb_group_new(devs, 3, buf, 512)
    assert g, buf.value.decode()
    try:
        for seed, n in ((1, 150_000), (1, 150_000), (2, 90_000), (3, 220_000)):
            sym = synth.repeats_dna(n, seed, unit=4000, copies=6, exact_len=3000, exact_copies=3)
            enc = encode_symbols(sym, 4)
            words, ranges = enc.twobitencoding(None)
            ranges = np.ascontiguousarray(ranges, dtype=np.uint64)
            assert lib.gtb_group_set_input_2bit(g, words.ctypes.data, words.shape[0], n,
                                                ranges.ctypes.data if ranges.shape[0] else None, ranges.shape[0]) == 0
            assert lib.gtb_group_run(g, 6, 7) == 0, lib.gtb_group_error(g).decode()
            h4 = (C.c_uint64 * 4)()
            assert lib.gtb_group_hash_results(g, h4) == 0
            ref = build_esa(enc, 6)
            im = images(ref)
            for i, (ext, dt) in enumerate((("suf", "<u8"), ("lcp", "u1"), ("llv", "<u8"), ("bck", "<u4"))):
                assert h4[i] == mixhash(np.frombuffer(im[ext], dtype=dt)), (seed, n, ext)
    finally:
        lib.gtb_group_delete(g)


@pytest.mark.parametrize("m,scan", [(16, "slice"), (20, "slice"), (24, "filter"), (20, "filter")])
def test_group_tail_keys_sorted_apart(monkeypatch, m, scan):
    """the first-level sort without the tail pass inside a sharded job -- the position-sharded scan (the
    owner splits the tail keys off the pairs it received) and the filtered scan of the whole text --
    forced on inputs full of specials"""
    monkeypatch.setenv("GTB200_KEY_SYMBOLS", str(m))
    monkeypatch.setenv("GTB200_TAIL_LAST", "1")
    monkeypatch.setenv("GTB200_SHARD_SCAN", scan)
    cases = [("reads", synth.reads(700, 60, 7, p_n=0.01), 4),
             ("repeats", synth.repeats_dna(90_000, 5, unit=2500, copies=7, exact_len=1200, exact_copies=3, nruns=6), 5),
             ("polyT_N", np.concatenate([np.full(3000, 3, np.uint8), synth.random_dna(2000, 3, p_n=0.02),
                                         np.full(2500, 3, np.uint8), [254], np.full(700, 3, np.uint8)]).astype(np.uint8), 3)]
    for name, sym, pl in cases:
        o = eo.esa(sym, 4, pl)
        im = eo.file_images(o)
        for nranges in (2, 3, 5):
            got = images(build_esa(encode_symbols(sym, 4), pl, devices=[0] * nranges))
            for ext in ("bck", "suf", "lcp", "llv"):
                assert got[ext] == im[ext], (name, m, scan, nranges, ext)
