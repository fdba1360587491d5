"""Wall time of the drop-in binary on a 64 Mbp FASTA sample of c4, stage by stage (its -v lines), with the
library's FASTA encoder and with the reference's (GTB200_ENCODER=reference), and the reference binary's
encoder alone (`gtref suffixerator -dna -tis`).  Run on the GPU box: python tools/cli_probe.py"""
import os, sys, subprocess, time, tempfile
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from genometools_b200 import synthetic as sy
import synth
w = sy.make_workload("c4", 64_000_000 / 3_100_000_000)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "s.fa"); synth.to_fasta(w.to_symbols(), fa, "dna")
print("host cores", os.cpu_count(), "fasta bytes", os.path.getsize(fa))
exe = os.path.join(ROOT, "host", "_build", "gt_b200")
variants = (("library encoder", {}), ("library encoder, output files not prefilled", {"GTB200_PREFILL": "0"}),
            ("library encoder, quick exit", {"GTB200_QUICK_EXIT": "1"}),
            ("reference encoder", {"GTB200_ENCODER": "reference"}))
if len(sys.argv) > 1:
    variants = variants[:int(sys.argv[1])]
for label, env in variants:
    for i in range(3):
        w0 = time.time()
        t0 = time.perf_counter()
        r = subprocess.run([exe, "suffixerator", "-dna", "-suf", "-lcp", "-bck", "-pl", "-v", "-indexname",
                            os.path.join(tmp, "x"), "-db", fa], capture_output=True, text=True,
                           env=dict(os.environ, GTB200_TRACE_WALL="1", **env))
        w1 = time.time()
        st = {l.split()[1]: float(l.split()[2]) for l in r.stderr.split("\n") if l.startswith("wallstamp")}
        where = ("before main %.3f, init %.3f, tool %.3f, cleanup %.3f, after main %.3f" %
                 (st["main"] - w0, st["tool"] - st["main"], st["tool_done"] - st["tool"],
                  st["main_done"] - st["tool_done"], w1 - st["main_done"])) if len(st) == 4 else str(st)
        print(label, round(time.perf_counter() - t0, 3), where,
              [l for l in r.stdout.split("\n") if "wall seconds" in l or "B200 encoder" in l])
gtref = os.path.join(ROOT, "oracle", "_ref", "gtref")
if os.path.exists(gtref):
    for i in range(2):
        t0 = time.perf_counter()
        subprocess.run([gtref, "suffixerator", "-dna", "-tis", "-indexname", os.path.join(tmp, "r"), "-db", fa],
                       stdout=subprocess.DEVNULL)
        print("gtref -tis only", round(time.perf_counter() - t0, 3))
    for i in range(2):
        t0 = time.perf_counter()
        subprocess.run([exe, "suffixerator", "-dna", "-tis", "-indexname", os.path.join(tmp, "o"), "-db", fa],
                       stdout=subprocess.DEVNULL)
        print("gt_b200 -tis only", round(time.perf_counter() - t0, 3))
    same = all(subprocess.call(["cmp", "-s", os.path.join(tmp, "r." + e), os.path.join(tmp, "o." + e)]) == 0
               for e in ("esq", "des", "sds", "md5", "prj"))
    print("index files identical:", same)
