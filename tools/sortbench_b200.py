#!/usr/bin/env python
"""`gt dev sortbench -impl radixinplace|radixkeypair` beside the B200 record sorts
(SURVEY.md section 8f row 4): Mkeys/s through the host-buffer C-ABI (H2D + sort + D2H inside the
timed region) and of the unmodified reference functions (oracle/_ref/gtref radixsort, one host core).
    python tools/sortbench_b200.py [records]          # prints one JSON line per kind
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genometools_b200 import _lib      # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    nref = min(n, 20_000_000)
    lib = _lib.load()
    gtref = os.path.join(ROOT, "oracle", "_ref", "gtref")
    rng = np.random.default_rng(9)
    for kind, width, fn in (("ulong", 1, lib.gtb_radixsort_u64), ("ulongpair", 2, lib.gtb_radixsort_u64pair),
                            ("keypair", 2, lib.gtb_radixsort_u64keypair)):
        a = rng.integers(0, 2 ** 63, size=(n, width), dtype=np.uint64)
        buf = ctypes.create_string_buffer(256)
        best = None
        for _ in range(3):
            b = a.copy()
            t0 = time.perf_counter()
            assert fn(0, b.ctypes.data, n, buf, 256) == 0, buf.value
            t = time.perf_counter() - t0
            best = t if best is None else min(best, t)
        assert (np.diff(b[:, 0].astype(np.float64)) >= 0).all()
        ref = None
        if os.path.exists(gtref):
            with tempfile.TemporaryDirectory() as tmp:
                a[:nref].tofile(os.path.join(tmp, "in"))
                ref = float(subprocess.check_output([gtref, "radixsort", kind, os.path.join(tmp, "in"),
                                                     os.path.join(tmp, "out")]).decode())
        print(json.dumps({"kind": kind, "records": n, "b200_seconds_host_to_host": best,
                          "b200_mrecords_per_s": n / best / 1e6,
                          "reference_records": nref, "reference_seconds_1_core": ref,
                          "reference_mrecords_per_s": (nref / ref / 1e6) if ref else None}))


if __name__ == "__main__":
    main()
