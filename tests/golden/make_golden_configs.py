"""Golden checksums of the BASELINE.json configurations AT FULL SIZE.

Runs the UNMODIFIED reference (oracle/_ref/gtref, built by `make -C oracle ref` from
/root/reference) once on each of c2, c3, c4 (-parts 4), c5 as genometools_b200/synthetic.py
generates them, and stores for every index file its length, md5 and the order-dependent
checksum of genometools_b200/mixhash.py (the one the GPU side computes over its shards in
HBM) in tests/golden/config_md5.json, together with the .prj text and a checksum of the
generated input (so that a drifting generator is noticed before a drifting sorter is blamed).

Run here (needs /root/reference and about 60 GB of scratch disk for c4):
    python tests/golden/make_golden_configs.py c2 c5 c3 c4      [--scale f] [--keep DIR]
bench.py and tests/test_gpu_configs.py only read the JSON.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from genometools_b200 import synthetic as sy          # noqa: E402
from genometools_b200.mixhash import mixhash, mixhash_file, FILE_DTYPES   # noqa: E402
import synth                                          # noqa: E402

OUT = os.path.join(HERE, "config_md5.json")
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 26), b""):
            h.update(blk)
    return h.hexdigest()


def input_checksum(w):
    if w.is_dna:
        return {"words": mixhash(w.words), "ranges": mixhash(w.ranges)}
    return {"symbols": mixhash(w.symbols)}


def key_of(name, scale):
    return name if scale == 1.0 else f"{name}@{scale:g}"


def run(name, scale, keep):
    t0 = time.time()
    w = sy.make_workload(name, scale)
    tmp = keep or tempfile.mkdtemp(prefix="gtb_cfg_", dir=os.environ.get("GTB_SCRATCH", "/tmp"))
    os.makedirs(tmp, exist_ok=True)
    try:
        fa = os.path.join(tmp, f"{name}.fa")
        synth.to_fasta(w.to_symbols(), fa, "dna" if w.is_dna else "protein")
        t1 = time.time()
        parts = ["-parts", "4"] if w.totallength > 2_000_000_000 else []
        cmd = [GTREF, "suffixerator", "-dna" if w.is_dna else "-protein", "-suf", "-lcp", "-bck", "-pl"] + parts + \
              ["-indexname", os.path.join(tmp, name), "-db", fa]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        t2 = time.time()
        files = {}
        for ext, dt in FILE_DTYPES.items():
            p = os.path.join(tmp, f"{name}.{ext}")
            files[ext] = {"bytes": os.path.getsize(p), "md5": md5_file(p), "mixhash": mixhash_file(p, dt)}
        prj = open(os.path.join(tmp, f"{name}.prj")).read()
        files["prj"] = {"bytes": len(prj), "md5": hashlib.md5(prj.encode()).hexdigest()}
        d = dict(line.split("=", 1) for line in prj.strip().split("\n"))
        rec = {"workload": name, "scale": scale, "description": w.description, "totallength": w.totallength,
               "numofchars": w.numofchars, "numofsequences": w.numofsequences,
               "specialcharacters": int(d["specialcharacters"]), "prefixlength": int(d["prefixlength"]),
               "command": " ".join(["gtref"] + cmd[1:-4] + ["-indexname", name, "-db", f"{name}.fa"]),
               "reference_seconds": round(t2 - t1, 1), "input": input_checksum(w), "files": files, "prj": prj}
        print(f"{key_of(name, scale)}: n={w.totallength} generated+fasta {t1 - t0:.0f}s, reference {t2 - t1:.0f}s, "
              f"checksums {time.time() - t2:.0f}s", flush=True)
        return rec
    finally:
        if not keep:
            shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--keep", default=None, help="directory to leave the FASTA and the index files in")
    args = ap.parse_args()
    if not os.path.exists(GTREF):
        sys.exit("oracle/_ref/gtref missing: run `make -C oracle -j8 ref` first")
    for name in args.configs:
        rec = run(name, args.scale, args.keep)
        table = json.load(open(OUT)) if os.path.exists(OUT) else {}
        table[key_of(name, args.scale)] = rec
        with open(OUT, "w") as fh:
            json.dump(table, fh, indent=1, sort_keys=True)
            fh.write("\n")


if __name__ == "__main__":
    main()
