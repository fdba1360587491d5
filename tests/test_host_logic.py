"""CPU: host-side mirror of the reference interface (no compute calls on the GPU)."""
import ctypes
import os
import re
import numpy as np
import pytest

import esa_oracle as eo
import synth
from conftest import ROOT, golden_cases
from genometools_b200 import _lib, encseq, sharding, suffixerator as sfx


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gtb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gtb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()                      # raises if the .so is missing
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gtb_abi_version() == 3


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.gtb_device_count() > 0:
        pytest.skip("CUDA device present")
    buf = ctypes.create_string_buffer(256)
    assert not lib.gtb_esa_new(0, buf, 256)
    assert b"no CPU fallback" in buf.value
    with pytest.raises(_lib.GtbError):
        sfx.Suffixerator(0)
    k = np.arange(4, dtype=np.uint64)
    v = np.arange(4, dtype=np.uint32)
    assert lib.gtb_radixsort_pairs_u64_u32(0, k.ctypes.data, v.ctypes.data, 4, 0, 64, buf, 256) == -1
    for fn in (lib.gtb_radixsort_u64, lib.gtb_radixsort_u64pair, lib.gtb_radixsort_u64keypair):
        buf2 = ctypes.create_string_buffer(256)
        assert fn(0, k.ctypes.data, 2, buf2, 256) == -1 and b"no CPU fallback" in buf2.value


@pytest.mark.parametrize("case", golden_cases())
def test_prefixlength_policy_matches_reference(golden, case):
    if not case.endswith("/auto") and not case.startswith("synth/"):
        return
    m = golden.meta(case)
    if case.startswith("synth/") and synth.SYNTH_CASES[case.split("/")[1]][3] is not None:
        return
    n = golden.symbols(case).shape[0] if golden.has(case, "symbols") else int(golden.prj(case)[0]["totallength"])
    assert sfx.recommendedprefixlength(m["numofchars"], n) == m["prefixlength"]


def test_maximal_prefixlength_known_answers():
    # the reference rejects -pl 7 for Atinsert.fna: "maximal prefix length ... is 6"
    assert sfx.whatisthemaximalprefixlength(4, 11817) == 6
    assert sfx.maxbasepower(4) == 15 and sfx.maxbasepower(20) == 7
    # c2..c5
    assert [sfx.recommendedprefixlength(4, n) for n in (100_000_000, 1_509_999_999, 3_100_000_000)] == [11, 13, 13]
    assert sfx.recommendedprefixlength(20, 500_000_000) == 5
    assert sfx.bcktab_sizeoftable(4, 11, 10 ** 8 + 1) == 22369620     # .bck of c2 is 22 369 624 with padding


@pytest.mark.parametrize("name", ["Atinsert.fna", "RandomN.fna", "TTTN.fna", "Duplicate.fna"])
def test_fasta_encoder_matches_golden_symbols(golden, name, tmp_path):
    case = f"file/{name}/auto"
    sym = golden.symbols(case)
    fa = tmp_path / name
    synth.to_fasta(sym, str(fa), "dna")
    enc = encseq.encode_fasta(str(fa), "dna")
    assert np.array_equal(enc.symbols, sym)
    assert enc.numofsequences == golden.meta(case)["numofsequences"]
    s2, nseq2 = eo.read_fasta(str(fa), "dna")
    assert np.array_equal(s2, sym) and nseq2 == enc.numofsequences
    prj, _ = golden.prj(case)
    info = enc.specialcharinfo()
    for k in ("specialcharacters", "realspecialranges", "wildcards", "realwildcardranges",
              "lengthofspecialprefix", "lengthofspecialsuffix"):
        assert info[k] == int(prj[k]), k


def test_fasta_errors():
    with pytest.raises(ValueError):
        encseq.parse_fasta_bytes(b"ACGT\n", encseq.ALPHABETS["dna"][1])
    with pytest.raises(ValueError):
        encseq.parse_fasta_bytes(b">x\nACGJ\n", encseq.ALPHABETS["dna"][1])
    s, n = encseq.parse_fasta_bytes(b">a desc\nAC\nGT\n>b\nNNa\n", encseq.ALPHABETS["dna"][1])
    assert s.tolist() == [0, 1, 2, 3, 255, 254, 254, 0] and n == 2


def test_twobit_export_layout():
    rng = np.random.default_rng(3)
    sym = rng.integers(0, 4, size=1000, dtype=np.uint8)
    sym[[5, 6, 7, 64, 999]] = [254, 254, 255, 254, 255]
    enc = encseq.encode_symbols(sym, 4)
    words, ranges = enc.twobitencoding()
    assert words.dtype == np.uint64 and words.shape[0] == 1000 // 32 + 2
    for i in (0, 1, 31, 32, 33, 500, 998):
        assert (int(words[i // 32]) >> (62 - 2 * (i % 32))) & 3 == sym[i]       # intbits.h:78-83
    assert ranges.tolist() == [[5, 8], [64, 65], [999, 1000]]
    w2, _ = enc.twobitencoding(filler=3)
    assert (int(w2[0]) >> (62 - 2 * 5)) & 3 == 3


def test_bck_and_prj_serialisation(golden):
    case = "file/Atinsert.fna/auto"
    o = eo.esa(golden.symbols(case), 4, 4)
    r = sfx.EsaResult(11817, 4, 4, leftborder=o["leftborder"].astype(np.uint32),
                      countspecialcodes=o["countspecialcodes"].astype(np.uint32),
                      distpfxidx=o["distpfxidx"].astype(np.uint32), longest=int(o["longest"]),
                      numoflargelcpvalues=0, maxbranchdepth=int(o["maxbranchdepth"]), lcptabsum=o["lcptabsum"])
    assert r.bck_bytes() == bytes(golden.get(case, "bck"))
    enc = encseq.encode_symbols(golden.symbols(case), 4, 21)
    assert r.prj_text(enc.specialcharinfo(), 21) == golden.prj(case)[1]


def test_parts_cover_all_codes():
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 50, size=4 ** 5)
    cnt[rng.random(cnt.size) < 0.3] = 0
    lb = np.concatenate(([0], np.cumsum(cnt)))
    for parts in (1, 2, 3, 8, 100):
        pl = sharding.suftab_parts(lb, parts)
        assert 1 <= len(pl) <= parts
        assert pl[0][0] == 0 and pl[-1][1] == 4 ** 5 - 1
        assert sum(p[3] for p in pl) == lb[-1]
        for a, b in zip(pl, pl[1:]):
            assert b[0] == a[1] + 1 and b[2] == a[2] + a[3]
        if parts > 1 and len(pl) == parts:
            assert max(p[3] for p in pl) <= lb[-1] // parts + cnt.max() + 1


def test_option_parser_mirrors_reference_errors():
    P = sfx.SuffixeratorOptions.parse
    o = P(["-dna", "-suf", "-lcp", "-bck", "-pl", "-db", "x.fna", "-indexname", "at"])
    assert o.pl == 0 and o.suf and o.lcp and o.bck and o.db == ["x.fna"]
    assert P(["-db", "a", "-dna", "-pl", "5", "-parts", "3"]).pl == 5
    for bad in (["-dna"], ["-db", "a", "b", "-dna"], ["-db", "a", "-dna", "-protein"],
                ["-db", "a", "-dir", "rev"], ["-db", "a", "-dc", "32"], ["-db", "a", "-bogus"],
                ["-db", "a", "-dna", "-suf", "-dir", "up"], ["-db", "a", "-protein", "-suf", "-dir", "rcl"]):
        with pytest.raises(_lib.GtbError):
            P(bad)
    # -dir (src/core/readmode.c:25-46, sfx-run.c:541-549,586-593)
    assert P(["-db", "a", "-dna", "-suf", "-dir", "rcl"]).dir == "rcl"
    assert P(["-db", "a", "-protein", "-lcp", "-dir", "rev"]).dir == "rev"
    with pytest.raises(_lib.GtbError, match="only makes sense"):
        P(["-db", "a", "-dna", "-bck", "-dir", "rev"])
    with pytest.raises(_lib.GtbError, match="only can be used for DNA"):
        P(["-db", "a", "-protein", "-suf", "-dir", "cpl"])
