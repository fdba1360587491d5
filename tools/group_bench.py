#!/usr/bin/env python
"""One job on the GPUs of this box inside ONE process (gtb_group: what `gt_b200 -j N` runs):
device time and wall time per step, per-range statistics, checksums against the reference.
    python tools/group_bench.py [--workload c4] [--scale 1.0] [--gpus 2] [--parts 1] [--steps 3]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genometools_b200 import _lib, synthetic as sy                      # noqa: E402
from genometools_b200._lib import GtbStats, ptr                         # noqa: E402
from genometools_b200.suffixerator import recommendedprefixlength       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--parts", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--copy", action="store_true", help="also time the result gather into one host table")
    args = ap.parse_args()
    lib = _lib.load()
    w = sy.make_workload(args.workload, args.scale)
    n = w.totallength
    pl = recommendedprefixlength(w.numofchars, n)
    devs = [i % args.gpus for i in range(args.gpus * args.parts)]
    buf = C.create_string_buffer(512)
    g = lib.gtb_group_new((C.c_int * len(devs))(*devs), len(devs), buf, 512)
    if not g:
        raise SystemExit(buf.value.decode())

    def ck(rc):
        if rc != 0:
            raise SystemExit(lib.gtb_group_error(g).decode())
    t0 = time.perf_counter()
    if w.is_dna:
        ck(lib.gtb_group_set_input_2bit(g, ptr(w.words), w.words.shape[0], n, ptr(w.ranges) if w.ranges.shape[0] else None,
                                        w.ranges.shape[0]))
    else:
        ck(lib.gtb_group_set_input_bytes(g, ptr(w.symbols), n, w.numofchars))
    t_up = time.perf_counter() - t0
    walls = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ck(lib.gtb_group_run(g, pl, 7))
        walls.append(time.perf_counter() - t0)
    st = GtbStats()
    ck(lib.gtb_group_get_stats(g, C.byref(st)))
    h4 = (C.c_uint64 * 4)()
    ck(lib.gtb_group_hash_results(g, h4))
    out = {"workload": args.workload, "scale": args.scale, "totallength": n, "devices": devs, "upload_s": t_up,
           "wall_ms_per_step": [x * 1e3 for x in walls], "Msuffixes_per_s": (n + 1) / min(walls) / 1e6,
           "job": {k: v for k, v in st.as_dict().items() if k.startswith("ms_") or k in ("doubling_rounds", "radix_passes", "kernel_launches")},
           "hashes": {"suf": h4[0], "lcp": h4[1], "llv": h4[2], "bck": h4[3]}}
    key = args.workload if args.scale == 1.0 else f"{args.workload}@{args.scale:g}"
    gp = os.path.join(ROOT, "tests", "golden", "config_md5.json")
    gold = json.load(open(gp)).get(key) if os.path.exists(gp) else None
    if gold:
        out["identical_to_reference"] = all(out["hashes"][e] == gold["files"][e]["mixhash"] for e in ("suf", "lcp", "llv", "bck"))
    ranges = []
    for i in range(len(devs)):
        s = GtbStats()
        lib.gtb_esa_get_stats(lib.gtb_group_range(g, i), C.byref(s))
        d = s.as_dict()
        ranges.append({k: d[k] for k in ("nonspecials", "sa_offset", "unresolved_after_first_sort", "ms_total", "ms_count", "ms_hist",
                                         "ms_radix", "ms_analyze", "ms_doubling", "ms_lcp")})
    out["ranges"] = ranges
    if args.copy:
        import numpy as np
        import torch
        e = n + 1
        suf = torch.empty(e, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        lcp = torch.empty(e, dtype=torch.uint8, pin_memory=True).numpy()
        k = int(lib.gtb_group_num_llv(g))
        llv = np.empty(2 * max(k, 1), dtype=np.uint64)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            ck(lib.gtb_group_copy_results(g, ptr(suf), ptr(lcp), ptr(llv) if k else None, None, None, None))
            ts.append(time.perf_counter() - t0)
        out["gather_ms"] = [x * 1e3 for x in ts]
    print(json.dumps(out))
    lib.gtb_group_delete(g)


if __name__ == "__main__":
    main()
