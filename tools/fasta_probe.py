"""Phase times of gtb_fasta_encode on a 64 Mbp FASTA for 1..all threads and the ways of mapping the input.
Run on the GPU box's host: python tools/fasta_probe.py"""
import os, sys, tempfile, json
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from genometools_b200 import synthetic as sy
from genometools_b200.encseq import write_index_files
import synth
w = sy.make_workload("c4", 64_000_000 / 3_100_000_000)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "s.fa"); synth.to_fasta(w.to_symbols(), fa, "dna")
print("host cores", os.cpu_count())
for how in ("plain", "populate", "seq"):
    os.environ["GTB200_FASTA_MAP"] = how
    for t in (1, 2, 4, 8, 16, 0):
        best = None
        for rep in range(3):
            s = write_index_files(fa, os.path.join(tmp, "o"), threads=t)
            if best is None or s["seconds_total"] < best["seconds_total"]:
                best = s
        print(how, "threads", best["threads"], " ".join("%s %.3f" % (k[8:], best[k]) for k in
              ("seconds_total", "seconds_count", "seconds_emit", "seconds_pack", "seconds_md5", "seconds_write")))
