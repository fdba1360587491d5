"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which runs the
reference on them) and by the parity tests (which regenerate them from the seed).
Shapes follow SURVEY.md section 8d / BASELINE.json configs 2-5 at test scale."""
import numpy as np

WILDCARD, SEPARATOR = 254, 255


def random_dna(n, seed, p_n=0.0):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 4, size=n, dtype=np.uint8)
    if p_n > 0:
        s[rng.random(n) < p_n] = WILDCARD
    return s


def reads(nreads, length, seed, p_n=0.001):
    """config 3: reads with N wildcards, one separator between reads"""
    rng = np.random.default_rng(seed)
    body = rng.integers(0, 4, size=(nreads, length), dtype=np.uint8)
    body[rng.random((nreads, length)) < p_n] = WILDCARD
    full = np.full((nreads, length + 1), SEPARATOR, dtype=np.uint8)
    full[:, :length] = body
    return np.ascontiguousarray(full.reshape(-1)[:-1])


def repeats_dna(n, seed, unit=3000, copies=6, exact_len=1500, exact_copies=3, nruns=3, nrun_len=40,
                mut=0.01):
    """config 4 at test scale: uniform background, mutated copies of a unit, exact
    duplicates (lcp >= 255 -> .llv) and runs of N"""
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 4, size=n, dtype=np.uint8)
    u = rng.integers(0, 4, size=unit, dtype=np.uint8)
    for _ in range(copies):
        c = u.copy()
        m = rng.random(unit) < mut
        c[m] = (c[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        p = int(rng.integers(0, n - unit))
        s[p:p + unit] = c
    e = rng.integers(0, 4, size=exact_len, dtype=np.uint8)
    for _ in range(exact_copies):
        p = int(rng.integers(0, n - exact_len))
        s[p:p + exact_len] = e
    for _ in range(nruns):
        p = int(rng.integers(0, n - nrun_len))
        s[p:p + nrun_len] = WILDCARD
    return s


def low_complexity_dna(n, seed):
    """poly-A / short tandem repeats: one giant bucket, very deep lcps"""
    rng = np.random.default_rng(seed)
    s = np.zeros(n, dtype=np.uint8)
    k = n // 3
    s[k:2 * k] = np.tile(np.array([0, 1], dtype=np.uint8), k)[:k]
    s[2 * k:] = rng.integers(0, 2, size=n - 2 * k, dtype=np.uint8)
    s[n // 2] = WILDCARD
    return s


def protein(nres, seed, reclen=350, p_x=0.001):
    """config 5: protein records (20 letters), X wildcards, separators"""
    rng = np.random.default_rng(seed)
    freq = np.array([9.9, 6.9, 5.9, 3.9, 5.8, 5.5, 6.7, 5.5, 8.3, 7.1, 6.6, 5.3, 4.1, 3.9, 2.9, 1.1, 4.7, 2.3,
                     2.4, 1.4])
    s = rng.choice(20, size=nres, p=freq / freq.sum()).astype(np.uint8)
    s[rng.random(nres) < p_x] = WILDCARD
    s[reclen::reclen + 1] = SEPARATOR
    if s[-1] == SEPARATOR:
        s[-1] = 0
    return s


def to_fasta(symbols, path, alphabet="dna"):
    letters = np.frombuffer(("ACGT" if alphabet == "dna" else "LVIFKREDAGSTNQYWPHMC").encode(), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    lut[: letters.size] = letters
    lut[WILDCARD] = ord("N" if alphabet == "dna" else "X")
    sym = np.asarray(symbols, dtype=np.uint8)
    seps = np.flatnonzero(sym == SEPARATOR)
    bounds = np.concatenate(([-1], seps, [sym.size]))
    with open(path, "wb") as fh:
        for i in range(len(bounds) - 1):
            rec = sym[bounds[i] + 1: bounds[i + 1]]
            fh.write(b">s%d\n" % i)
            fh.write(lut[rec].tobytes())
            fh.write(b"\n")


# name -> (generator, alphabet, numofchars, prefixlength or None=auto)
SYNTH_CASES = {
    "rand_dna_50k": (lambda: random_dna(50_000, 42), "dna", 4, None),
    "rand_dna_N_60k": (lambda: random_dna(60_000, 43, p_n=0.002), "dna", 4, None),
    "reads_400x100": (lambda: reads(400, 100, 1), "dna", 4, None),
    "reads_dup_300x80": (lambda: np.concatenate([reads(300, 80, 2, 0.0), [SEPARATOR], reads(300, 80, 2, 0.0)]).astype(np.uint8), "dna", 4, 5),
    "repeats_80k": (lambda: repeats_dna(80_000, 7), "dna", 4, None),
    "lowcomplex_30k": (lambda: low_complexity_dna(30_000, 5), "dna", 4, 3),
    "protein_40k": (lambda: protein(40_000, 11), "protein", 20, None),
    "protein_dup_20k": (lambda: np.concatenate([protein(10_000, 12), [SEPARATOR], protein(10_000, 12)]).astype(np.uint8), "protein", 20, 2),
    "rand_dna_2M": (lambda: random_dna(2_000_000, 44, p_n=0.0001), "dna", 4, None),
}
BIG_CASES = {"rand_dna_2M"}     # golden holds md5 sums only
