import os, sys, subprocess, time, tempfile
ROOT=os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests","golden"))
from genometools_b200 import synthetic as sy
import synth
w = sy.make_workload("c4", 64_000_000/3_100_000_000)
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "s.fa"); synth.to_fasta(w.to_symbols(), fa, "dna")
for i in range(3):
    t0=time.perf_counter()
    r=subprocess.run([os.path.join(ROOT,"host","_build","gt_b200"),"suffixerator","-dna","-suf","-lcp","-bck","-pl","-v","-indexname",os.path.join(tmp,"x"),"-db",fa],capture_output=True,text=True)
    print(round(time.perf_counter()-t0,3), [l for l in r.stdout.split("\n") if "wall seconds" in l or "B200" in l], r.stderr[-200:])
