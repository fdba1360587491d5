"""CPU: the FASTA -> encseq encoder of libgtb200.so (gtb_fasta_encode, host code)
against the unmodified reference.

tests/golden/fasta_index_md5.json holds the md5 of every index file `gtref suffixerator -dna -tis
[-des/-sds/-ssp/-md5 yes|no] [-clipdesc]` wrote for the inputs of tests/golden/fasta_cases.py
(make_golden_fasta.py; the reference's encoder is gt_encseq_new_from_files,
/root/reference/src/core/encseq.c:7503-7714; `-protein` for the protein cases).  The library must write the
same bytes: .esq (header, 2-bit words, wildcard range table or special bits; 5-bit string for protein), .ssp,
.des, .sds, .md5 -- with one chunk per file
and with chunk borders every 50 bytes (the parallel decomposition must not show).  Where the reference
binary is present (this container) the drop-in binary is run beside it on inputs both accept and
inputs the library declines.
"""
import hashlib
import json
import os
import subprocess

import pytest

import encseq_oracle
import fasta_cases
from conftest import ROOT
from genometools_b200.encseq import FastaUnsupported, write_index_files

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "fasta_index_md5.json")))
CASES = fasta_cases.all_cases()
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")
GT_B200 = os.path.join(ROOT, "host", "_build", "gt_b200")
SUFFIXES = ("esq", "ssp", "des", "sds", "md5")


def write_inputs(files, d, opts=None):
    names = []
    for name, data in fasta_cases.file_names_and_bytes(files, opts or {}):
        names.append(name)                  # relative: the .esq header stores the names as given
        (d / name).write_bytes(data)
    return names


def check_case(name, tmp_path, monkeypatch, chunk=None, threads=0):
    files, opts = CASES[name]
    g = GOLDEN[name]
    assert [hashlib.md5(f).hexdigest() for f in files] == g["input_md5"], "generator drifted from the golden inputs"
    monkeypatch.chdir(tmp_path)
    if chunk:
        monkeypatch.setenv("GTB200_FASTA_CHUNK", str(chunk))
    names = write_inputs(files, tmp_path, opts)
    s = write_index_files(names, "our", threads=threads, **{k: v for k, v in opts.items() if k != "gz"})
    for suf in SUFFIXES:
        p = tmp_path / ("our." + suf)
        assert p.exists() == (suf in g["files"]), f"{name}: .{suf} written: {p.exists()}, reference: {suf in g['files']}"
        if p.exists():
            data = p.read_bytes()
            assert len(data) == g["files"][suf]["bytes"], f"{name}: .{suf} length"
            assert hashlib.md5(data).hexdigest() == g["files"][suf]["md5"], f"{name}: .{suf} differs from the reference's"
    prj = dict(line.split("=", 1) for line in g["prj"].strip().split("\n"))
    for key in ("totallength", "specialcharacters", "specialranges", "realspecialranges", "wildcards",
                "wildcardranges", "realwildcardranges"):
        assert s[key] == int(prj[key]), key
    assert s["numofsequences"] == int(prj["numofdbsequences"])
    return s


@pytest.mark.parametrize("name", sorted(CASES))
def test_index_files_identical_to_reference(name, tmp_path, monkeypatch):
    check_case(name, tmp_path, monkeypatch)


@pytest.mark.parametrize("chunk", [50, 7])
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith(("small_0", "odd_", "protein_0"))])
def test_chunk_borders_do_not_show(name, chunk, tmp_path, monkeypatch):
    """the files are cut every `chunk` bytes (anywhere but inside a description): in the middle of lines, of
    wildcard runs, of CR LF pairs, right in front of a '>'"""
    check_case(name, tmp_path, monkeypatch, chunk=chunk, threads=3)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith("odd_")])
def test_a_chunk_per_character(name, tmp_path, monkeypatch):
    check_case(name, tmp_path, monkeypatch, chunk=1, threads=2)


def test_sequence_on_one_line_is_cut_too(tmp_path, monkeypatch):
    s = check_case("one_line_3M", tmp_path, monkeypatch, chunk=1 << 16, threads=4)
    assert s["totallength"] == 3_000_000


def test_every_representation_is_reached(tmp_path, monkeypatch):
    seen = set()
    for i, name in enumerate(("uint32_single", "ushort_multi", "equal_length_reads", "small_004", "small_000")):
        d = tmp_path / str(i)
        d.mkdir()
        seen.add(check_case(name, d, monkeypatch)["satname"])
    for name in sorted(CASES):
        if len(seen) == 5:
            break
        if name.startswith("small_"):
            d = tmp_path / name
            d.mkdir()
            seen.add(check_case(name, d, monkeypatch)["satname"])
    assert seen == {"eqlen", "bit", "uchar", "ushort", "uint32"}, seen


def test_lower_case_protein_is_declined(tmp_path, monkeypatch):
    """the protein alphabet maps upper case only (assignproteinsymbolmap, src/core/alphabet.c:488-506)"""
    monkeypatch.chdir(tmp_path)
    (tmp_path / "p.fa").write_bytes(b">p\nMKVlaag\n")
    with pytest.raises(FastaUnsupported):
        write_index_files(["p.fa"], "our", alphabet="protein")
    assert sorted(os.listdir(tmp_path)) == ["p.fa"]


DECLINED = {
    "illegal_character": [b">a\nACGT\nACXT\n"],
    "empty_sequence": [b">a\nACGT\n>b\n>c\nAC\n"],
    "empty_last_sequence": [b">a\nACGT\n>b\n"],
    "description_cut_off": [b">a\nACGT\n>b"],
    "no_header": [b"ACGT\n"],
    "second_file_without_header": [b">a\nACGT\n", b"ACGT\n"],
    "nul_in_description": [b">a\0b\nACGT\n"],
    "empty_file": [b""],
}


@pytest.mark.parametrize("name", sorted(DECLINED))
def test_declined_inputs_write_nothing(name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    names = write_inputs(DECLINED[name], tmp_path)
    with pytest.raises(FastaUnsupported):
        write_index_files(names, "our")
    assert sorted(os.listdir(tmp_path)) == sorted(names)


def test_broken_and_missing_files_are_declined(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "a.fa.bz2").write_bytes(b">a\nACGT\n")         # libbz2 passes nothing through that is not bzip2
    with pytest.raises(FastaUnsupported):
        write_index_files(["a.fa.bz2"], "our")
    (tmp_path / "broken.fa.gz").write_bytes(b"\x1f\x8b\x08\x00 this is not a deflate stream")
    with pytest.raises(FastaUnsupported):
        write_index_files(["broken.fa.gz"], "our")
    with pytest.raises(FastaUnsupported):
        write_index_files(["missing.fa"], "our")


def test_index_directory_that_does_not_exist_is_declined(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "a.fa").write_bytes(b">a\nACGT\n")
    with pytest.raises(FastaUnsupported):
        write_index_files(["a.fa"], "no/such/dir/our")


def run_tool(exe, args, cwd):
    return subprocess.run([exe, "suffixerator"] + args, cwd=cwd, capture_output=True, text=True)


needs_binaries = pytest.mark.skipif(not (os.path.exists(GTREF) and os.path.exists(GT_B200)),
                                    reason="oracle/_ref/gtref and host/_build/gt_b200 are built where /root/reference is")


@needs_binaries
@pytest.mark.parametrize("name,extra", [
    ("three_files", ["-dna"]), ("three_files", []), ("many_sequences", ["-dna", "-clipdesc"]),
    ("odd_header_in_midline", ["-dna", "-md5", "no"]), ("small_004", ["-dna", "-des", "no", "-sds", "no"]),
    ("ushort_multi", ["-dna", "-sat", "uint32"]), ("small_001", ["-dna", "-lossless"]),
    ("protein_1M", ["-protein"]), ("protein_03", []), ("protein_05", ["-protein", "-sat", "direct"]),
    ("gz_mixed_three_files", ["-dna"]), ("gz_protein_04", []),
])
def test_dropin_binary_without_sort(name, extra, tmp_path):
    """`gt_b200 suffixerator -tis` (no table requested: no GPU involved) beside `gtref`: same files, the
    .prj included; -sat and -lossless go through the reference's encoder inside the drop-in"""
    files, opts = CASES[name]
    out = {}
    for who, exe in (("ref", GTREF), ("our", GT_B200)):
        d = tmp_path / who
        d.mkdir()
        names = write_inputs(files, d, opts)
        r = run_tool(exe, ["-tis", "-v", "-indexname", "i", "-db"] + names + extra, d)
        assert r.returncode == 0, r.stderr
        out[who] = ({f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.startswith("i.")}, r.stdout)
    assert sorted(out["ref"][0]) == sorted(out["our"][0])
    for f in out["ref"][0]:
        assert out["ref"][0][f] == out["our"][0][f], f
    fast = "B200 encoder:" in out["our"][1]
    assert fast == ("-lossless" not in extra)


@needs_binaries
def test_dropin_binary_when_the_index_cannot_be_written(tmp_path):
    res = {}
    for who, exe in (("ref", GTREF), ("our", GT_B200)):
        d = tmp_path / who
        d.mkdir()
        names = write_inputs([b">a\nACGT\n"], d)
        r = run_tool(exe, ["-dna", "-tis", "-indexname", "no/such/dir/i", "-db"] + names, d)
        res[who] = (r.returncode, r.stderr.split(": error: ", 1)[-1], sorted(os.listdir(d)))
    assert res["ref"][0] != 0
    assert res["ref"] == res["our"]


@needs_binaries
@pytest.mark.parametrize("name", sorted(DECLINED))
def test_dropin_binary_on_declined_inputs(name, tmp_path):
    """what the library declines the reference's encoder handles inside the drop-in: same exit code, same
    message, same files left behind"""
    res = {}
    for who, exe in (("ref", GTREF), ("our", GT_B200)):
        d = tmp_path / who
        d.mkdir()
        names = write_inputs(DECLINED[name], d)
        r = run_tool(exe, ["-dna", "-tis", "-indexname", "i", "-db"] + names, d)
        msg = r.stderr.split(": error: ", 1)[-1]
        res[who] = (r.returncode, msg, {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.startswith("i.")})
    assert res["ref"][0] == res["our"][0]
    assert res["ref"][1] == res["our"][1]
    assert res["ref"][2] == res["our"][2]


TESTDATA = "/root/reference/testdata"


@needs_binaries
@pytest.mark.skipif(not os.path.isdir(TESTDATA), reason="the reference's testdata is only in this container")
def test_reference_testdata_through_dropin(tmp_path):
    """every FASTA-like file of the reference's testdata (alphabet guessed, as in
    testsuite/gt_suffixerator_include.rb) through `suffixerator -tis` of both binaries: same exit code, same
    message, same files -- whether the library encodes it (DNA) or declines it (protein, malformed input)"""
    import glob
    files = []
    for f in sorted(glob.glob(os.path.join(TESTDATA, "*"))):
        if os.path.isfile(f) and 0 < os.path.getsize(f) < 5_000_000:
            with open(f, "rb") as fh:
                if fh.read(1) == b">":
                    files.append(f)
    assert len(files) > 100
    fast = 0
    for f in files:
        res = {}
        for who, exe in (("ref", GTREF), ("our", GT_B200)):
            d = tmp_path / who
            if d.exists():
                for x in os.listdir(d):
                    os.unlink(d / x)
            else:
                d.mkdir()
            r = run_tool(exe, ["-tis", "-v", "-indexname", str(d / "i"), "-db", f], tmp_path)
            res[who] = (r.returncode, r.stderr.split(": error: ", 1)[-1],
                        {x: hashlib.md5((d / x).read_bytes()).hexdigest() for x in sorted(os.listdir(d))})
            if who == "our" and "B200 encoder:" in r.stdout:
                fast += 1
        assert res["ref"] == res["our"], f
    assert fast > 100          # most of them are DNA


ENCODER_BIN = os.path.join(ROOT, "host", "_build", "gtref_b200_encoder")


@pytest.mark.skipif(not (os.path.exists(GTREF) and os.path.exists(ENCODER_BIN)),
                    reason="oracle/_ref/gtref and host/_build/gtref_b200_encoder are built where /root/reference is")
@pytest.mark.parametrize("name,tool,extra,fast", [
    ("ushort_multi", "suffixerator", ["-dna", "-suf", "-lcp", "-bck", "-pl"], True),
    ("three_files", "suffixerator", ["-tis"], True),
    ("protein_1M", "suffixerator", ["-protein", "-tis"], True),
    ("many_sequences", "packedindex_mkindex", ["-dna", "-tis"], True),
    ("small_004", "suffixerator", ["-dna", "-tis", "-clipdesc", "-md5", "no"], True),
    ("small_001", "suffixerator", ["-dna", "-tis", "-lossless"], False),
    ("ushort_multi", "suffixerator", ["-dna", "-tis", "-sat", "uint32"], True),
])
def test_encoder_api_under_the_reference_tools(name, tool, extra, fast, tmp_path):
    """host/gt_encseq_encoder_b200.c = gt_encseq_encoder_encode (src/core/encseq_api.h:347) on top of the library,
    linked into the UNMODIFIED reference tools (its own suffixerator with the CPU sorter, packedindex mkindex): every
    file they write is the file the all-reference binary writes; -lossless / -sat reach the reference's function"""
    files, _ = CASES[name]
    out = {}
    for who, exe in (("ref", GTREF), ("our", ENCODER_BIN)):
        d = tmp_path / who
        d.mkdir()
        names = write_inputs(files, d)
        r = subprocess.run([exe, tool, "-indexname", "i", "-db"] + names + extra, cwd=d, capture_output=True, text=True,
                           env=dict(os.environ, GTB200_TRACE_ENCODER="1"))
        assert r.returncode == 0, r.stderr
        out[who] = ({f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.startswith("i.")}, r.stderr)
    assert sorted(out["ref"][0]) == sorted(out["our"][0])
    for f in out["ref"][0]:
        assert out["ref"][0][f] == out["our"][0][f], f
    assert ("B200 encoder:" in out["our"][1]) == fast


@pytest.mark.skipif(not (os.path.exists(GTREF) and os.path.exists(ENCODER_BIN)),
                    reason="oracle/_ref/gtref and host/_build/gtref_b200_encoder are built where /root/reference is")
@pytest.mark.parametrize("name", sorted(DECLINED))
def test_encoder_api_on_declined_inputs(name, tmp_path):
    res = {}
    for who, exe in (("ref", GTREF), ("our", ENCODER_BIN)):
        d = tmp_path / who
        d.mkdir()
        names = write_inputs(DECLINED[name], d)
        r = subprocess.run([exe, "suffixerator", "-dna", "-tis", "-indexname", "i", "-db"] + names, cwd=d,
                           capture_output=True, text=True)
        res[who] = (r.returncode, r.stderr.split(": error: ", 1)[-1],
                    {f: (d / f).read_bytes() for f in sorted(os.listdir(d)) if f.startswith("i.")})
    assert res["ref"] == res["our"]


# ---- the sequential restatement (oracle/encseq_oracle.py): pinned to the reference, then used as the checker

ORACLE_CASES = [n for n in sorted(CASES) if sum(len(f) for f in CASES[n][0]) < 200_000]


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_matches_reference(name):
    files, opts = CASES[name]
    g = GOLDEN[name]
    names = [nm for nm, _ in fasta_cases.file_names_and_bytes(files, opts)]
    out = encseq_oracle.encode(files, names, **{k: v for k, v in opts.items() if k != "gz"})
    assert sorted(out) == sorted(g["files"])
    for suf, data in out.items():
        assert hashlib.md5(data).hexdigest() == g["files"][suf]["md5"], suf


def random_fasta(rng, alphabet):
    letters = "ACGT" if alphabet == "dna" else fasta_cases.AMINO
    wild = fasta_cases.WILD if alphabet == "dna" else fasta_cases.AMINO_WILD
    out = []
    for s in range(rng.choice([1, 2, 5, 30])):
        n = rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 700, 3000])
        q = [rng.choice(letters) for _ in range(n)]
        for _ in range(rng.choice([0, 0, 1, 3, 40])):
            a = rng.randrange(n)
            for i in range(a, min(n, a + rng.choice([1, 2, 9, 300]))):
                q[i] = rng.choice(wild)
        eol = rng.choice(["\n", "\r\n"])
        width = rng.choice([1, 13, 60, 10 ** 9])
        text = eol.join("".join(q[i:i + width]) for i in range(0, n, width))
        if rng.random() < 0.3:
            k = rng.randrange(len(text) + 1)
            text = text[:k] + rng.choice([" ", "\t", eol + eol]) + text[k:]
        out.append(">%s%s%s" % (rng.choice(["", "id%d" % s, "id %d\tx  y" % s]), eol, text) +
                   (eol if rng.random() < 0.9 or s + 1 < 30 else ""))
    return "".join(out).encode()


@pytest.mark.parametrize("seed", range(60))
def test_library_against_oracle_on_random_inputs(seed, tmp_path, monkeypatch):
    """inputs that are in no golden file: random option sets, widths, page-border lengths; random chunk size"""
    import random
    rng = random.Random(5000 + seed)
    alphabet = rng.choice(["dna", "dna", "protein"])
    sat = rng.choice([None, None] + (["uchar", "ushort", "uint32", "bit", "direct", "eqlen"] if alphabet == "dna"
                                     else ["bytecompress", "direct", "ushort"]))
    files = [random_fasta(rng, alphabet) for _ in range(rng.choice([1, 1, 2, 3]))]
    opts = {k: rng.random() < 0.7 for k in ("des", "ssp", "md5")}
    opts["sds"] = opts["des"] and rng.random() < 0.7
    opts["clip_desc"] = rng.random() < 0.3
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("GTB200_FASTA_CHUNK", str(rng.choice([1, 5, 64, 1000, 1 << 20])))
    names = write_inputs(files, tmp_path)
    try:
        want = encseq_oracle.encode(files, names, alphabet=alphabet, sat=sat, **opts)
    except encseq_oracle.Declined:
        with pytest.raises(FastaUnsupported):
            write_index_files(names, "our", alphabet=alphabet, sat=sat, **opts)
        return
    write_index_files(names, "our", alphabet=alphabet, sat=sat, threads=rng.choice([1, 2, 5]), **opts)
    for suf in SUFFIXES:
        p = tmp_path / ("our." + suf)
        assert p.exists() == (suf in want), suf
        if p.exists():
            assert p.read_bytes() == want[suf], suf
