"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/gtref, built by `make -C oracle ref` from /root/reference) on

  * the FASTA fixtures the reference's own suffixerator tests use
    (/root/reference/testsuite/gt_suffixerator_include.rb:119-143, :292-309),
    with automatic and explicit prefix lengths,
  * the seeded synthetic inputs of tests/golden/synth.py.

Run here (needs /root/reference); the parity tests only read the .npz files.
    python tests/golden/make_golden.py
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

TESTDATA = "/root/reference/testdata"
DNA_FILES = ["Arabidopsis-C99826.fna", "Atinsert.fna", "Atinsert_seqrange_13-17_rev.fna",
             "Atinsert_seqrange_3-7.fna", "Atinsert_single_3.fna", "Atinsert_single_3_rev.fna",
             "Copysorttest.fna", "Duplicate.fna", "Ecoli-section1.fna", "Ecoli-section2.fna",
             "Random-Small.fna", "Random.fna", "Random159.fna", "Random160.fna", "RandomN.fna",
             "Reads1.fna", "Reads2.fna", "Reads3.fna", "Repfind-example.fna", "TTTN.fna", "Small.fna",
             "Smalldup.fna", "TTT-small.fna", "trna_glutamine.fna", "Verysmall.fna"]
PROTEIN_FILES = ["sw100K1.fsa", "sw100K2.fsa"]
MULTI = [("multi_RandomN_Random_Atinsert", ["RandomN.fna", "Random.fna", "Atinsert.fna"], "dna"),
         ("multi_sw100K1_sw100K2", ["sw100K1.fsa", "sw100K2.fsa"], "protein")]


def prj_dict(text):
    return dict(line.split("=", 1) for line in text.decode().strip().split("\n"))


def pack(case, symbols, nseq, alphabet, K, ref, full=True):
    prj = prj_dict(ref["prj"])
    d = {"symbols": symbols, "numofsequences": np.int64(nseq), "numofchars": np.int64(K),
         "alphabet": alphabet, "prefixlength": np.int64(int(prj["prefixlength"])),
         "prj": np.frombuffer(ref["prj"], dtype=np.uint8)}
    for ext in ("suf", "lcp", "llv", "bck"):
        d["md5_" + ext] = hashlib.md5(ref[ext]).hexdigest()
        d["len_" + ext] = np.int64(len(ref[ext]))
    if full:
        d["suf"] = np.frombuffer(ref["suf"], dtype="<u8").astype(np.uint32)
        d["lcp"] = np.frombuffer(ref["lcp"], dtype=np.uint8)
        d["llv"] = np.frombuffer(ref["llv"], dtype="<u8")
        d["bck"] = np.frombuffer(ref["bck"], dtype=np.uint8)
    return {f"{case}/{k}": v for k, v in d.items()}


def main():
    if not eo.have_reference():
        sys.exit("oracle/_ref/gtref missing: run `make -C oracle -j8 ref` first")
    out, names = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        def add(case, paths, alphabet, pl, symbols=None, nseq=None, full=True):
            K = 4 if alphabet == "dna" else 20
            ref = eo.run_reference(paths, tmp, alphabet, pl)
            ref3 = eo.run_reference(paths, tmp, alphabet, pl, parts=3, indexname="ref3")
            for ext in ("suf", "lcp", "llv", "bck"):
                assert ref[ext] == ref3[ext], (case, ext, "-parts 3 changed the output")
            if symbols is None:
                symbols, nseq = eo.read_fasta(paths, alphabet)
            prj = prj_dict(ref["prj"])
            assert int(prj["totallength"]) == symbols.shape[0], case
            assert int(prj["numofsequences"]) == nseq, case
            out.update(pack(case, symbols, nseq, alphabet, K, ref, full))
            names.append(case)
            print(f"{case:45s} n={symbols.shape[0]:8d} pl={prj['prefixlength']:>2s} "
                  f"maxbranchdepth={prj['maxbranchdepth']} large={prj['largelcpvalues']}")

        for f in DNA_FILES:
            p = os.path.join(TESTDATA, f)
            add(f"file/{f}/auto", p, "dna", None)
        for f, pl in [("Atinsert.fna", 1), ("Atinsert.fna", 2), ("Atinsert.fna", 3), ("Atinsert.fna", 6),
                      ("RandomN.fna", 2), ("RandomN.fna", 6), ("Duplicate.fna", 3), ("Reads2.fna", 5),
                      ("TTTN.fna", 1), ("Random.fna", 6)]:
            add(f"file/{f}/pl{pl}", os.path.join(TESTDATA, f), "dna", pl)
        for f in PROTEIN_FILES:
            add(f"file/{f}/auto", os.path.join(TESTDATA, f), "protein", None)
            add(f"file/{f}/pl2", os.path.join(TESTDATA, f), "protein", 2)
        for name, files, alpha in MULTI:
            add(f"{name}/auto", [os.path.join(TESTDATA, f) for f in files], alpha, None)
        add("multi_RandomN_Random_Atinsert/pl5", [os.path.join(TESTDATA, f) for f in MULTI[0][1]], "dna", 5)
        for name, (gen, alpha, K, pl) in synth.SYNTH_CASES.items():
            sym = gen()
            fa = os.path.join(tmp, name + ".fa")
            synth.to_fasta(sym, fa, alpha)
            nseq = int((sym == 255).sum()) + 1
            add(f"synth/{name}", fa, alpha, pl, symbols=sym, nseq=nseq, full=name not in synth.BIG_CASES)
    big = {k: v for k, v in out.items() if k.split("/")[1] in synth.BIG_CASES and k.endswith("/symbols")}
    for k in big:
        del out[k]              # regenerated from the seed at test time
    out["__cases__"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(names), "cases")


if __name__ == "__main__":
    main()
