"""Seeded synthetic inputs of the five BASELINE.json configurations (SURVEY.md
section 8d), produced directly in the form the C-ABI takes (2-bit words + special
ranges, or symbol bytes) so that a 3.1 Gbp input is built in seconds.

  c2  100 Mbp uniform DNA, one sequence                      (symbol level, seed 42)
  c3  reads of 150 bp with N wildcards and separators        (symbol level, seed 1)
  c4  human-sized DNA with planted long repeats and N runs   (word level,  seed 7)
  c5  protein records, 20 letters, X wildcards               (symbol level, seed 11)
"""
from dataclasses import dataclass
import numpy as np

WILDCARD, SEPARATOR = 254, 255


@dataclass
class Workload:
    name: str
    totallength: int
    numofchars: int
    numofsequences: int
    words: np.ndarray = None       # uint64 2-bit words (DNA)
    ranges: np.ndarray = None      # uint64 [r,2] special ranges (DNA)
    symbols: np.ndarray = None     # uint8 (byte path, or kept for small DNA inputs)
    description: str = ""

    @property
    def is_dna(self):
        return self.numofchars == 4

    @property
    def specialcharacters(self):
        if self.ranges is not None:
            return int((self.ranges[:, 1] - self.ranges[:, 0]).sum())
        return int((self.symbols >= WILDCARD).sum())

    def input_bytes(self):
        if self.is_dna:
            return int(self.words.nbytes + self.ranges.nbytes)
        return int(self.symbols.nbytes)

    def to_symbols(self):
        """symbol bytes (for the CPU reference / oracle on bounded samples)"""
        if self.symbols is not None:
            return self.symbols
        n = self.totallength
        b = self.words.astype(">u8").view(np.uint8)
        s = np.empty(b.size * 4, dtype=np.uint8)
        s[0::4] = b >> 6; s[1::4] = (b >> 4) & 3; s[2::4] = (b >> 2) & 3; s[3::4] = b & 3
        s = s[:n].copy()
        for a, e in self.ranges:
            s[int(a):int(e)] = WILDCARD
        return s


def pack_symbols(sym):
    from .encseq import encode_symbols
    enc = encode_symbols(sym, 4)
    return enc.twobitencoding()


def c2_uniform_dna(n=100_000_000, seed=42):
    rng = np.random.default_rng(seed)
    sym = rng.integers(0, 4, size=n, dtype=np.uint8)
    words, ranges = pack_symbols(sym)
    return Workload("c2", n, 4, 1, words, ranges, sym if n <= 50_000_000 else None,
                    f"uniform random DNA, one sequence of {n} bp, seed {seed}")


def c3_reads(nreads=10_000_000, length=150, seed=1, p_n=0.001):
    rng = np.random.default_rng(seed)
    n = nreads * (length + 1) - 1
    sym = np.empty(n + 1, dtype=np.uint8)
    chunk = 1_000_000
    for a in range(0, nreads, chunk):
        b = min(nreads, a + chunk)
        body = rng.integers(0, 4, size=(b - a, length + 1), dtype=np.uint8)
        body[:, :length][rng.random((b - a, length)) < p_n] = WILDCARD
        body[:, length] = SEPARATOR
        sym[a * (length + 1): b * (length + 1)] = body.reshape(-1)
    sym = sym[:n]
    words, ranges = pack_symbols(sym)
    return Workload("c3", n, 4, nreads, words, ranges, sym if n <= 50_000_000 else None,
                    f"{nreads} reads x {length} bp, p(N)={p_n}, separators, seed {seed}")


def c4_repeat_dna(n=3_100_000_000, seed=7, scale=None):
    """uniform background + planted repeats (word aligned: positions and lengths are
    multiples of 32 bp): 2000 x 6 kbp, 200 x 50 kbp, 20 x 1 Mbp copies mutated at 1 %,
    5 exact 100 kbp duplicates, 0.5 % N in runs of 1000.  `scale` < 1 builds a
    bounded sample with the same share of the text in every repeat family."""
    rng = np.random.default_rng(seed)
    if scale is None:
        scale = n / 3_100_000_000
    nw = n // 32 + 2
    words = rng.integers(0, 2 ** 63, size=nw, dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=nw, dtype=np.uint64)
    usable = n // 32 - 1

    def plant(unit_bp, copies, mut):
        # bounded samples keep the share of the text each family covers: fewer copies
        # (at least 2) and, when that is not enough, shorter units
        covered = unit_bp * copies * scale
        ncopies = max(2, int(round(copies * scale)))
        unit_len = int(min(unit_bp, covered / ncopies)) // 32 * 32
        L = unit_len // 32
        if L < 2 or L >= usable // 4:
            return
        unit = rng.integers(0, 2 ** 63, size=L, dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=L, dtype=np.uint64)
        for _ in range(ncopies):
            w = unit.copy()
            if mut > 0:
                k = int(unit_len * mut)
                pos = rng.integers(0, unit_len, size=k)
                delta = rng.integers(1, 4, size=k).astype(np.uint64)
                np.bitwise_xor.at(w, pos // 32, delta << ((31 - pos % 32) * 2).astype(np.uint64))
            p = int(rng.integers(0, usable - L))
            words[p:p + L] = w

    plant(6_016, 2000, 0.01)
    plant(50_016, 200, 0.01)
    plant(1_000_000 // 32 * 32, 20, 0.01)
    plant(100_000 // 32 * 32, 5, 0.0)
    # N runs of 1000 on a jittered grid (sorted, disjoint)
    nruns = int(n * 0.005 / 1000)
    if nruns > 0:
        step = n // nruns
        starts = (np.arange(nruns, dtype=np.int64) * step + rng.integers(0, max(1, step - 1001), size=nruns))
        ranges = np.stack([starts, starts + 1000], axis=1).astype(np.uint64)
        ranges = ranges[ranges[:, 1] <= n]
    else:
        ranges = np.zeros((0, 2), dtype=np.uint64)
    # bases beyond n are zero (as in the exported encoding)
    rem = n % 32
    if rem:
        words[n // 32] &= np.uint64(((1 << (2 * rem)) - 1) << (64 - 2 * rem))
    else:
        words[n // 32] = 0
    words[n // 32 + 1:] = 0
    return Workload("c4", n, 4, 1, words, np.ascontiguousarray(ranges), None,
                    f"{n} bp uniform DNA + planted repeats (2000x6k, 200x50k, 20x1M at 1%, 5 exact 100k; "
                    f"repeat families scaled x{scale:.3g}) + 0.5% N in runs of 1000, seed {seed}")


def c5_protein(nres=500_000_000, seed=11, reclen=350, p_x=0.0001):
    rng = np.random.default_rng(seed)
    freq = np.array([9.9, 6.9, 5.9, 3.9, 5.8, 5.5, 6.7, 5.5, 8.3, 7.1, 6.6, 5.3, 4.1, 3.9, 2.9, 1.1, 4.7, 2.3,
                     2.4, 1.4])
    cdf = np.cumsum(freq / freq.sum())
    sym = np.empty(nres, dtype=np.uint8)
    chunk = 50_000_000
    for a in range(0, nres, chunk):
        b = min(nres, a + chunk)
        r = rng.random(b - a)
        sym[a:b] = np.minimum(np.searchsorted(cdf, r), 19).astype(np.uint8)
        sym[a:b][rng.random(b - a) < p_x] = WILDCARD
    sym[reclen::reclen + 1] = SEPARATOR
    if sym[-1] >= WILDCARD:
        sym[-1] = 0
    nseq = int((sym == SEPARATOR).sum()) + 1
    return Workload("c5", nres, 20, nseq, None, None, sym,
                    f"{nres} residues in records of {reclen}, Swiss-Prot-like background, p(X)={p_x}, seed {seed}")


def make_workload(name, scale=1.0):
    if name == "c2":
        return c2_uniform_dna(int(100_000_000 * scale))
    if name == "c3":
        return c3_reads(int(10_000_000 * scale))
    if name == "c4":
        return c4_repeat_dna(int(3_100_000_000 * scale), scale=scale)
    if name == "c5":
        return c5_protein(int(500_000_000 * scale))
    raise ValueError(f"unknown workload {name}")
