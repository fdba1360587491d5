"""Generate tests/golden/bwt_vectors.npz: the `.bwt` files the UNMODIFIED reference
(oracle/_ref/gtref suffixerator ... -bwt) writes for a subset of the cases of
reference_vectors.npz (first "next" row).  Needs /root/reference.
    python tests/golden/make_golden_bwt.py
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
from make_golden import TESTDATA, DNA_FILES, PROTEIN_FILES, MULTI   # noqa: E402


def main():
    if not eo.have_reference():
        sys.exit("oracle/_ref/gtref missing: run `make -C oracle -j8 ref` first")
    out, names = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        def add(case, paths, alphabet, pl):
            ref = eo.run_reference(paths, tmp, alphabet, pl, extra=("-bwt",))
            ref3 = eo.run_reference(paths, tmp, alphabet, pl, parts=3, indexname="ref3", extra=("-bwt",))
            assert ref["bwt"] == ref3["bwt"], (case, "-parts 3 changed the .bwt")
            out[case + "/bwt"] = np.frombuffer(ref["bwt"], dtype=np.uint8)
            out[case + "/md5_bwt"] = hashlib.md5(ref["bwt"]).hexdigest()
            out[case + "/md5_suf"] = hashlib.md5(ref["suf"]).hexdigest()    # ties the vector to its case
            names.append(case)
            print(f"{case:45s} {len(ref['bwt']):9d} bytes")

        for f in DNA_FILES:
            add(f"file/{f}/auto", os.path.join(TESTDATA, f), "dna", None)
        for f in PROTEIN_FILES:
            add(f"file/{f}/auto", os.path.join(TESTDATA, f), "protein", None)
        for name, files, alpha in MULTI:
            add(f"{name}/auto", [os.path.join(TESTDATA, f) for f in files], alpha, None)
        for name, (gen, alpha, K, pl) in synth.SYNTH_CASES.items():
            if name in synth.BIG_CASES:
                continue
            fa = os.path.join(tmp, name + ".fa")
            synth.to_fasta(gen(), fa, alpha)
            add(f"synth/{name}", fa, alpha, pl)
    out["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "bwt_vectors.npz"), **out)
    print("wrote bwt_vectors.npz", len(names), "cases")


if __name__ == "__main__":
    main()
