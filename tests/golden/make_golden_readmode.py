"""Generate tests/golden/readmode_vectors.npz: what the UNMODIFIED reference
(oracle/_ref/gtref suffixerator ... -dir rev|cpl|rcl -suf -lcp -bck -bwt) writes for a subset of
the cases of reference_vectors.npz (first "next" row: reverse /
complement read modes, /root/reference/src/match/sfx-mapped4.gen:33-86,
src/core/encseq.c:6094-6140).  Needs /root/reference.
    python tests/golden/make_golden_readmode.py
"""
import hashlib
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import esa_oracle as eo     # noqa: E402
import synth                # noqa: E402
from make_golden import TESTDATA, prj_dict   # noqa: E402

DNA_FILES = ["Atinsert.fna", "RandomN.fna", "TTTN.fna", "Duplicate.fna", "Reads2.fna", "Random159.fna",
             "Random160.fna", "Verysmall.fna", "trna_glutamine.fna", "Ecoli-section1.fna"]
PROTEIN_FILES = ["sw100K1.fsa"]
SYNTH = ["reads_400x100", "repeats_80k", "rand_dna_N_60k", "protein_dup_20k"]


def main():
    if not eo.have_reference():
        sys.exit("oracle/_ref/gtref missing: run `make -C oracle -j8 ref` first")
    out, names = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        def add(case, paths, alphabet, pl, mode):
            ref = eo.run_reference(paths, tmp, alphabet, pl, extra=("-dir", mode, "-bwt"))
            ref3 = eo.run_reference(paths, tmp, alphabet, pl, parts=3, indexname="ref3", extra=("-dir", mode, "-bwt"))
            for ext in ("suf", "lcp", "llv", "bck", "bwt"):
                assert ref[ext] == ref3[ext], (case, ext, "-parts 3 changed the output")
            prj = prj_dict(ref["prj"])
            assert int(prj["readmode"]) == ("fwd", "rev", "cpl", "rcl").index(mode)
            key = f"{case}@{mode}"
            for ext in ("suf", "lcp", "llv", "bck", "bwt"):
                out[f"{key}/md5_{ext}"] = hashlib.md5(ref[ext]).hexdigest()
                out[f"{key}/len_{ext}"] = np.int64(len(ref[ext]))
            out[f"{key}/prj"] = np.frombuffer(ref["prj"], dtype=np.uint8)
            out[f"{key}/prefixlength"] = np.int64(int(prj["prefixlength"]))
            names.append(key)
            print(f"{key:45s} pl={prj['prefixlength']:>2s} longest={prj['longest']} maxbranchdepth={prj['maxbranchdepth']}")

        for f in DNA_FILES:
            for mode in ("rev", "cpl", "rcl"):
                add(f"file/{f}/auto", os.path.join(TESTDATA, f), "dna", None, mode)
        for f in PROTEIN_FILES:
            add(f"file/{f}/auto", os.path.join(TESTDATA, f), "protein", None, "rev")
        for name in SYNTH:
            gen, alpha, K, pl = synth.SYNTH_CASES[name]
            fa = os.path.join(tmp, name + ".fa")
            synth.to_fasta(gen(), fa, alpha)
            for mode in (("rev", "cpl", "rcl") if alpha == "dna" else ("rev",)):
                add(f"synth/{name}", fa, alpha, pl, mode)
    out["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "readmode_vectors.npz"), **out)
    print("wrote readmode_vectors.npz", len(names), "cases")


if __name__ == "__main__":
    main()
