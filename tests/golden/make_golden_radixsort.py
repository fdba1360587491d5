"""Generate tests/golden/radixsort_vectors.npz: md5 sums of what the UNMODIFIED reference's
gt_radixsort_inplace_ulong / _GtUwordPair / _Gtuint64keyPair / _flba (/root/reference/src/core/radix_sort.h:91,
107,125,138; called through `oracle/_ref/gtref radixsort`, oracle/ref_driver.c) return for seeded inputs
("next" row 4).  Needs /root/reference.
    python tests/golden/make_golden_radixsort.py
"""
import hashlib
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")

CASES = {  # name: (kind, records, seed, distinct values of the key or 0 = all 64 bits)
    "ulong_100k": ("ulong", 100_000, 1, 0), "ulong_dups": ("ulong", 50_000, 2, 300), "ulong_1": ("ulong", 1, 3, 0),
    "ulong_small": ("ulong", 37, 4, 0), "pair_100k": ("ulongpair", 100_000, 5, 0),
    "pair_dups": ("ulongpair", 60_000, 6, 1000), "keypair_100k": ("keypair", 100_000, 7, 0),
    "keypair_dups": ("keypair", 80_000, 8, 50),
}


# gt_radixsort_inplace_flba (radix_sort.h:138): records of `unitsize` bytes in memcmp order.
# name: (unitsize, records, seed, distinct byte values per position or 0 = all 256)
# (one-byte records: the reference itself dies with SIGSEGV on 10 000 of them; 37 are sorted by its insertion sort)
FLBA_CASES = {
    "flba1_small": (1, 37, 11, 0), "flba2": (2, 50_000, 20, 0), "flba3": (3, 100_000, 12, 0), "flba5_dups": (5, 100_000, 13, 3),
    "flba8": (8, 100_000, 14, 0), "flba9": (9, 100_000, 15, 0), "flba12_dups": (12, 80_000, 16, 2),
    "flba16": (16, 100_000, 17, 0), "flba7_small": (7, 37, 18, 0), "flba16_one": (16, 1, 19, 0),
}


def make_flba(unitsize, n, seed, distinct):
    rng = np.random.default_rng(seed)
    return rng.integers(0, distinct if distinct else 256, size=(n, unitsize), dtype=np.uint8)


def make_input(kind, n, seed, distinct):
    rng = np.random.default_rng(seed)
    w = 1 if kind == "ulong" else 2
    a = rng.integers(0, 2 ** 63, size=(n, w), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, w), dtype=np.uint64)
    if distinct:
        pool = rng.integers(0, 2 ** 63, size=distinct, dtype=np.uint64) * np.uint64(2)
        a[:, 0] = pool[rng.integers(0, distinct, size=n)]
        if kind == "keypair":
            a[:, 1] = pool[rng.integers(0, distinct, size=n)]
    return a


def main():
    if not os.path.exists(GTREF):
        sys.exit("oracle/_ref/gtref missing: run `make -C oracle -j8 ref` first")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (kind, n, seed, distinct) in CASES.items():
            a = make_input(kind, n, seed, distinct)
            a.tofile(os.path.join(tmp, "in"))
            subprocess.check_output([GTREF, "radixsort", kind, os.path.join(tmp, "in"), os.path.join(tmp, "out")])
            b = np.fromfile(os.path.join(tmp, "out"), dtype=np.uint64).reshape(a.shape)
            out[name + "/md5"] = hashlib.md5(b.tobytes()).hexdigest()
            out[name + "/md5_keys"] = hashlib.md5(np.ascontiguousarray(b[:, 0]).tobytes()).hexdigest()
            print(name, out[name + "/md5"])
        for name, (unitsize, n, seed, distinct) in FLBA_CASES.items():
            a = make_flba(unitsize, n, seed, distinct)
            a.tofile(os.path.join(tmp, "in"))
            subprocess.check_output([GTREF, "radixsort", "flba%d" % unitsize, os.path.join(tmp, "in"),
                                     os.path.join(tmp, "out")])
            b = np.fromfile(os.path.join(tmp, "out"), dtype=np.uint8).reshape(a.shape)
            out[name + "/md5"] = hashlib.md5(b.tobytes()).hexdigest()
            print(name, out[name + "/md5"])
    out["__cases__"] = np.array(list(CASES))
    np.savez_compressed(os.path.join(HERE, "radixsort_vectors.npz"), **out)


if __name__ == "__main__":
    main()
