"""GPU: the Sfxiterator-shaped object (host/gt_sfxiterator_b200.c).

host/_build/gt_b200_sfxiterator is the reference's own code -- its suffixerator driver
(src/match/sfx-run.c) and `gt packedindex mkindex` (src/tools/gt_packedindex.c:33-36 ->
src/match/eis-suffixerator-interface.c:253,399) -- linked with OUR definitions of
gt_Sfxiterator_new/_next/_longest/_bcktab2file/_delete in front of src/match/sfx-suffixer.o.
Everything those consumers write must equal what they write on top of the CPU iterator."""
import os
import subprocess

import numpy as np
import pytest

import synth
from conftest import ROOT

pytestmark = pytest.mark.gpu
GT_SFX = os.path.join(ROOT, "host", "_build", "gt_b200_sfxiterator")
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")
need_bins = pytest.mark.skipif(not (os.path.exists(GT_SFX) and os.path.exists(GTREF)),
                               reason="host/_build/gt_b200_sfxiterator or oracle/_ref/gtref not built")


def files_of(prefix):
    d, base = os.path.split(prefix)
    return {f[len(base) + 1:]: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d)) if f.startswith(base + ".")}


@need_bins
@pytest.mark.parametrize("case", ["reads", "repeats", "protein"])
def test_packedindex_mkindex_pulls_the_suffix_array_from_the_gpu(tmp_path, case):
    if case == "reads":
        sym, alpha = synth.reads(1500, 90, 4, 0.003), "dna"
    elif case == "repeats":
        sym, alpha = synth.repeats_dna(150_000, 8, unit=4000, copies=8, exact_len=12000, exact_copies=3), "dna"
    else:
        sym, alpha = synth.protein(40_000, 3), "protein"
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, alpha)
    out = {}
    for name, exe in (("b200", GT_SFX), ("ref", GTREF)):
        idx = str(tmp_path / name)
        r = subprocess.run([exe, "packedindex_mkindex", "-" + alpha, "-tis", "-v", "-db", fa, "-indexname", idx],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-600:]
        if name == "b200":
            assert "B200 Sfxiterator:" in r.stdout          # the GPU sorted, not the archive's CPU iterator
        out[name] = files_of(idx)
    assert sorted(out["b200"]) == sorted(out["ref"]) and "bdx" in out["ref"]
    for ext in out["ref"]:
        assert out["b200"][ext] == out["ref"][ext], ext


@need_bins
@pytest.mark.parametrize("parts", ["1", "3"])
def test_reference_suffixerator_driver_on_the_b200_iterator(tmp_path, parts):
    """the reference's OWN suffixeratorwithoutput() loop (sfx-run.c:212-317: next() until NULL, suftab pieces to
    file, longest, bcktab2file / flush at delete) driving our iterator: .suf .bck .prj as on the CPU iterator"""
    sym = np.concatenate([synth.reads(900, 80, 6, 0.004), np.array([255], dtype=np.uint8),
                          synth.repeats_dna(120_000, 3, unit=3000, copies=7, exact_len=8000, exact_copies=3)])
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(sym, fa, "dna")
    out = {}
    for name, exe in (("b200", GT_SFX), ("ref", GTREF)):
        idx = str(tmp_path / name)
        subprocess.check_call([exe, "suffixerator", "-dna", "-suf", "-bck", "-pl", "5", "-parts", parts, "-indexname", idx,
                               "-db", fa], stdout=subprocess.DEVNULL)
        out[name] = files_of(idx)
    for ext in ("suf", "bck", "prj", "esq"):
        assert out["b200"][ext] == out["ref"][ext], ext


@need_bins
def test_b200_iterator_refuses_what_it_does_not_do(tmp_path):
    fa = str(tmp_path / "in.fa")
    synth.to_fasta(synth.random_dna(20_000, 4, 0.001), fa, "dna")
    for extra, what in ((["-lcp"], "lcp side channel"), (["-dc", "32"], "difference cover")):
        r = subprocess.run([GT_SFX, "suffixerator", "-dna", "-suf", *extra, "-indexname", str(tmp_path / "x"), "-db", fa],
                           capture_output=True, text=True)
        assert r.returncode != 0 and "not supported by the B200 Sfxiterator" in r.stderr and what in r.stderr
