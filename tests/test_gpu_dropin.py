"""GPU: the C host drop-in (host/gt_suffixerator_b200.c linked with the reference's own
option parser, encoder and project-file writer -> host/_build/gt_b200) against the
unmodified reference binary (oracle/_ref/gtref) on the same FASTA files: all five index
files must be byte-identical, and the option set outside the path must fail loudly."""
import os
import subprocess
import numpy as np
import pytest

import synth
from conftest import ROOT

pytestmark = pytest.mark.gpu
GT_B200 = os.path.join(ROOT, "host", "_build", "gt_b200")
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")
need_bins = pytest.mark.skipif(not (os.path.exists(GT_B200) and os.path.exists(GTREF)),
                               reason="host/_build/gt_b200 or oracle/_ref/gtref not built")


def run_both(tmp_path, fasta, alphabet, extra=()):
    out = {}
    for name, exe in (("b200", GT_B200), ("ref", GTREF)):
        idx = str(tmp_path / name)
        subprocess.check_call([exe, "suffixerator", "-" + alphabet, "-suf", "-lcp", "-bck", "-pl", *extra,
                               "-indexname", idx, "-db", *fasta], stdout=subprocess.DEVNULL)
        out[name] = {ext: open(idx + "." + ext, "rb").read() for ext in ("suf", "lcp", "llv", "bck", "prj", "esq")}
        if "-bwt" in extra:
            out[name]["bwt"] = open(idx + ".bwt", "rb").read()
    return out


@need_bins
@pytest.mark.parametrize("case", ["repeats", "reads", "protein", "two_files"])
def test_dropin_files_identical_to_reference(tmp_path, case):
    if case == "repeats":
        sym, alpha = synth.repeats_dna(400_000, 31, unit=8000, copies=10, exact_len=30000, exact_copies=3), "dna"
    elif case == "reads":
        sym, alpha = synth.reads(4000, 100, 8, 0.002), "dna"
    elif case == "protein":
        sym, alpha = synth.protein(150_000, 21), "protein"
    else:
        sym, alpha = synth.random_dna(100_000, 9, 0.001), "dna"
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, alpha)
    files = [fa]
    if case == "two_files":
        fb = str(tmp_path / "in2.fa")
        synth.to_fasta(synth.reads(300, 90, 4), fb, alpha)
        files.append(fb)
    out = run_both(tmp_path, files, alpha)
    for ext in ("esq", "bck", "suf", "lcp", "llv", "prj"):
        assert out["b200"][ext] == out["ref"][ext], ext


@need_bins
@pytest.mark.parametrize("case", ["reads", "protein"])
def test_dropin_bwt_identical_to_reference(tmp_path, case):
    sym, alpha = (synth.reads(3000, 90, 12, 0.004), "dna") if case == "reads" else (synth.protein(60_000, 5), "protein")
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, alpha)
    # "-pl" takes the optional numeric argument, so the extra options follow an explicit value
    out = run_both(tmp_path, [fa], alpha, extra=("3" if alpha == "dna" else "2", "-bwt"))
    for ext in ("bwt", "bck", "suf", "lcp", "llv", "prj"):
        assert out["b200"][ext] == out["ref"][ext], ext


@need_bins
def test_dropin_explicit_prefixlength_and_errors(tmp_path):
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(synth.random_dna(50_000, 3, 0.01), fa, "dna")
    out = run_both(tmp_path, [fa], "dna", extra=("4",))
    for ext in ("bck", "suf", "lcp", "llv", "prj"):
        assert out["b200"][ext] == out["ref"][ext], ext
    for bad in (["-dc", "32"], ["-suftabuint"]):
        r = subprocess.run([GT_B200, "suffixerator", "-dna", "-suf", *bad, "-indexname", str(tmp_path / "x"), "-db", fa],
                           capture_output=True, text=True)
        assert r.returncode != 0 and "not supported by the B200" in r.stderr
    r = subprocess.run([GT_B200, "suffixerator", "-dna", "-suf", "-pl", "9", "-indexname", str(tmp_path / "x"), "-db", fa],
                       capture_output=True, text=True)
    r2 = subprocess.run([GTREF, "suffixerator", "-dna", "-suf", "-pl", "9", "-indexname", str(tmp_path / "y"), "-db", fa],
                        capture_output=True, text=True)
    assert r.returncode != 0 and r2.returncode != 0
    assert r.stderr.split("error:")[1] == r2.stderr.split("error:")[1]      # same message as the reference


@need_bins
@pytest.mark.parametrize("case,mode", [("reads", "rev"), ("reads", "cpl"), ("reads", "rcl"), ("repeats", "rcl"),
                                       ("protein", "rev")])
def test_dropin_readmodes_identical_to_reference(tmp_path, case, mode):
    """-dir rev|cpl|rcl: all files incl. .bwt and the readmode line of .prj"""
    if case == "reads":
        sym, alpha = synth.reads(2500, 100, 17, 0.003), "dna"
    elif case == "repeats":
        sym, alpha = synth.repeats_dna(200_000, 5, unit=5000, copies=8, exact_len=20000, exact_copies=3), "dna"
    else:
        sym, alpha = synth.protein(80_000, 9), "protein"
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, alpha)
    out = run_both(tmp_path, [fa], alpha, extra=("3" if alpha == "dna" else "2", "-bwt", "-dir", mode))
    for ext in ("bwt", "bck", "suf", "lcp", "llv", "prj"):
        assert out["b200"][ext] == out["ref"][ext], ext


@need_bins
def test_dropin_readmode_errors_like_the_reference(tmp_path):
    fa = str(tmp_path / "p.fa")
    synth.to_fasta(synth.protein(5_000, 2), fa, "protein")
    for args in (["-protein", "-suf", "-dir", "cpl"], ["-protein", "-dir", "rev"]):
        msgs = []
        for exe in (GT_B200, GTREF):
            r = subprocess.run([exe, "suffixerator", *args, "-indexname", str(tmp_path / "x"), "-db", fa],
                               capture_output=True, text=True)
            assert r.returncode != 0
            msgs.append(r.stderr.split("error:")[1])
        assert msgs[0] == msgs[1]


@need_bins
@pytest.mark.parametrize("mode", ["fwd", "rcl"])
def test_dropin_index_passes_the_reference_verifier(tmp_path, mode):
    """`gt dev sfxmap -suf -lcp -bck -esa` -- the reference's own brute-force checker
    (/root/reference/src/tools/gt_sfxmap.c, src/match/esa-map.c) reads OUR index files and accepts them"""
    sym = np.concatenate([synth.reads(400, 90, 3, 0.004), np.array([255], dtype=np.uint8),
                          synth.repeats_dna(60_000, 6, unit=3000, copies=6, exact_len=9000, exact_copies=3)])
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, "dna")
    idx = str(tmp_path / "b200")
    subprocess.check_call([GT_B200, "suffixerator", "-dna", "-tis", "-suf", "-lcp", "-bck", "-ssp", "-des", "-sds",
                           "-pl", "-dir", mode, "-indexname", idx, "-db", fa], stdout=subprocess.DEVNULL)
    r = subprocess.run([GTREF, "sfxmap", "-tis", "-suf", "-lcp", "-bck", "-ssp", "-v", "-esa", idx],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    assert "compare lcp-values against reference" in r.stdout and r.stdout.strip().endswith("okay")
    # and it does notice a wrong table: swap two suffix-table entries
    suf = np.fromfile(idx + ".suf", dtype=np.uint64)
    suf[[1000, 1001]] = suf[[1001, 1000]]
    suf.tofile(idx + ".suf")
    r = subprocess.run([GTREF, "sfxmap", "-tis", "-suf", "-lcp", "-bck", "-ssp", "-esa", idx], capture_output=True, text=True)
    assert r.returncode != 0


@need_bins
def test_dropin_rejects_mirrored_loudly(tmp_path):
    """-mirrored is outside the path (a mirrored GtEncseq reports 2n+1 symbols over n exported bases):
    non-zero exit, the library is never called, no index tables are written"""
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(synth.random_dna(20_000, 5, 0.001), fa, "dna")
    idx = str(tmp_path / "m")
    r = subprocess.run([GT_B200, "suffixerator", "-dna", "-suf", "-lcp", "-mirrored", "-indexname", idx, "-db", fa],
                       capture_output=True, text=True)
    assert r.returncode != 0
    assert "option -mirrored is not supported by the B200 suffixerator path" in r.stderr
    assert not os.path.exists(idx + ".suf") and not os.path.exists(idx + ".lcp")


@need_bins
def test_dropin_plain_input_like_the_reference(tmp_path):
    """-plain: the drop-in makes the reference's own calls (src/match/sfx-run.c:496-531, incl. the
    do-not-create-des/sds branch), so it behaves as the reference does -- in 1.5.11 the suffixerator's
    encoder is built from options that do not carry -plain and BOTH stop with the same message; should a
    reference build accept the input, the files must be identical"""
    raw = str(tmp_path / "in.txt")
    rng = np.random.default_rng(8)
    open(raw, "wb").write(bytes(rng.choice(np.frombuffer(b"acgt", dtype=np.uint8), size=30_000)))
    res = {}
    for name, exe in (("b200", GT_B200), ("ref", GTREF)):
        idx = str(tmp_path / name)
        r = subprocess.run([exe, "suffixerator", "-dna", "-plain", "-suf", "-lcp", "-bck", "-pl", "-indexname", idx,
                            "-db", raw], capture_output=True, text=True)
        res[name] = (r.returncode, r.stderr.split("error:")[-1] if r.returncode else "")
    assert res["b200"] == res["ref"]
    if res["ref"][0] == 0:
        for ext in ("suf", "lcp", "llv", "bck", "prj", "esq"):
            assert open(str(tmp_path / "b200") + "." + ext, "rb").read() == open(str(tmp_path / "ref") + "." + ext, "rb").read(), ext


@need_bins
@pytest.mark.parametrize("opts", [("-parts", "3"), ("-parts", "7")])
def test_dropin_parts_identical_to_reference(tmp_path, opts):
    """-parts p: p bucket-code ranges, the same files (testsuite/gt_suffixerator_include.rb:64-68)"""
    sym = np.concatenate([synth.reads(1500, 100, 5, 0.003), np.array([255], dtype=np.uint8),
                          synth.repeats_dna(250_000, 2, unit=6000, copies=9, exact_len=20000, exact_copies=3)])
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, "dna")
    out = run_both(tmp_path, [fa], "dna", extra=("6", "-bwt") + opts)
    for ext in ("bwt", "bck", "suf", "lcp", "llv", "prj"):
        assert out["b200"][ext] == out["ref"][ext], ext


def _ngpus():
    from genometools_b200 import _lib
    return _lib.load().gtb_device_count()


@need_bins
@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_dropin_multi_gpu_identical_to_reference(tmp_path, ngpu):
    """`gt -j N suffixerator`: one bucket-code range on each of N GPUs of the box, inside the drop-in
    binary; all five files byte-identical to the reference's"""
    if _ngpus() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    sym = np.concatenate([synth.reads(3000, 120, 15, 0.002), np.array([255], dtype=np.uint8),
                          synth.repeats_dna(1_500_000, 12, unit=30_000, copies=12, exact_len=120_000, exact_copies=4)])
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, "dna")
    ref = str(tmp_path / "ref")
    subprocess.check_call([GTREF, "suffixerator", "-dna", "-suf", "-lcp", "-bck", "-pl", "-bwt", "-indexname", ref,
                           "-db", fa], stdout=subprocess.DEVNULL)
    idx = str(tmp_path / "b200")
    r = subprocess.run([GT_B200, "-j", str(ngpu), "suffixerator", "-dna", "-suf", "-lcp", "-bck", "-pl", "-bwt", "-v",
                        "-indexname", idx, "-db", fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert f"B200: {ngpu} GPU(s), {ngpu} bucket-code range(s)" in r.stdout
    for ext in ("suf", "lcp", "llv", "bck", "bwt", "prj"):
        assert open(f"{idx}.{ext}", "rb").read() == open(f"{ref}.{ext}", "rb").read(), ext


@need_bins
def test_dropin_from_an_existing_index(tmp_path):
    """`suffixerator -ii <index>`: the sequences were encoded before (here by the drop-in itself, `-tis` only);
    both binaries build the tables from the loaded index (src/match/sfx-run.c:454-492)"""
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(synth.repeats_dna(200_000, 5, unit=3000, copies=6, exact_len=9000, exact_copies=3), fa, "dna")
    enc = str(tmp_path / "enc")
    subprocess.check_call([GT_B200, "suffixerator", "-dna", "-tis", "-indexname", enc, "-db", fa],
                          stdout=subprocess.DEVNULL)
    out = {}
    for name, exe in (("b200", GT_B200), ("ref", GTREF)):
        idx = str(tmp_path / name)
        subprocess.check_call([exe, "suffixerator", "-ii", enc, "-suf", "-lcp", "-bck", "-pl", "-indexname", idx],
                              stdout=subprocess.DEVNULL)
        out[name] = {ext: open(idx + "." + ext, "rb").read() for ext in ("suf", "lcp", "llv", "bck", "prj")}
        assert not os.path.exists(idx + ".esq")
    for ext in out["ref"]:
        assert out["b200"][ext] == out["ref"][ext], ext


RELEASE_SCRIPT = r"""
import ctypes as C, gc, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import synth
from genometools_b200 import _lib, encode_symbols
from genometools_b200.suffixerator import build_esa
lib = _lib.load()
sym = synth.random_dna(30_000, 11, 0.01)
first = build_esa(encode_symbols(sym, 4), prefixlength=4, device=0)
suf = first.suf_bytes()
buf = C.create_string_buffer(256)
h = lib.gtb_esa_new(0, buf, 256)
assert h
assert lib.gtb_release_devices() == 0, "a live handle: nothing may be released"
lib.gtb_esa_delete(h)
del first
gc.collect()
assert lib.gtb_release_devices() == 1, "no handle left on device 0: its context goes"
assert lib.gtb_release_devices() == 0, "nothing left to release"
again = build_esa(encode_symbols(sym, 4), prefixlength=4, device=0)
assert again.suf_bytes() == suf
print("ok")
"""


def test_release_devices_between_handles():
    """gtb_release_devices destroys the context of a device without live handles (what the drop-in does beside its
    file writes); the next handle builds a new one and sorts as before.  In a process of its own: the reset takes
    the context away from everything else in the process, torch included"""
    import sys
    r = subprocess.run([sys.executable, "-c", RELEASE_SCRIPT, ROOT], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
