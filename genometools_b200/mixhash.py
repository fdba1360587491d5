"""Order-dependent 64-bit checksum of an index table that composes across shards.

    H(table) = sum_i  fin( (i + 1) * C1  xor  fin(v_i + C2) )        (mod 2^64)

`fin` is the splitmix64 finaliser, i the GLOBAL index of entry v_i.  Because the index
takes part, two tables with the same multiset of values in a different order differ;
because the per-entry terms are summed, the checksum of a table is the sum of the
checksums of its shards (each computed with its global offset) -- every GPU hashes its
own shard in HBM (gtb_esa_hash_results, libgtb200.so) and the sums are added.

The same function over the reference's files (tests/golden/make_golden_configs.py runs
the unmodified reference at full configuration size and stores md5 + mixhash of
.suf/.lcp/.llv/.bck in tests/golden/config_md5.json) is what bench.py compares with at
every GPU count: equal checksums <=> identical tables, up to 2^-64 chance.

Entry conventions (must match the device side, csrc/gtb_esa.cu):
  .suf  uint64 entries, index = suffix-array index
  .lcp  uint8 entries,  index = suffix-array index
  .llv  the flat uint64 sequence of the file (index0, value0, index1, value1, ...)
  .bck  the uint32 words of the file, 8-byte padding words (value 0) included
"""
import numpy as np

C1 = np.uint64(0x9E3779B97F4A7C15)
C2 = np.uint64(0xC2B2AE3D27D4EB4F)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
MASK = (1 << 64) - 1


def _fin(z):
    z ^= z >> np.uint64(30)
    z *= M1
    z ^= z >> np.uint64(27)
    z *= M2
    z ^= z >> np.uint64(31)
    return z


def mixhash(values, start_index=0, chunk=1 << 24):
    """checksum of `values` (any unsigned integer array) whose first entry has global index
    `start_index`; returns a Python int < 2^64"""
    values = np.asarray(values).reshape(-1)
    total = 0
    with np.errstate(over="ignore"):
        for a in range(0, values.shape[0], chunk):
            v = values[a:a + chunk].astype(np.uint64)
            idx = np.arange(start_index + a + 1, start_index + a + 1 + v.shape[0], dtype=np.uint64)
            z = (idx * C1) ^ _fin(v + C2)
            total = (total + int(_fin(z).sum(dtype=np.uint64))) & MASK
    return total


def mixhash_file(path, dtype, chunk_bytes=1 << 28):
    """checksum of a raw little-endian file of `dtype` entries, read in chunks"""
    dt = np.dtype(dtype)
    total, index = 0, 0
    with open(path, "rb") as fh:
        while True:
            buf = fh.read(chunk_bytes)
            if not buf:
                break
            if len(buf) % dt.itemsize:
                raise ValueError(f"{path}: size is not a multiple of {dt.itemsize}")
            arr = np.frombuffer(buf, dtype=dt)
            total = (total + mixhash(arr, index)) & MASK
            index += arr.shape[0]
    return total


FILE_DTYPES = {"suf": "<u8", "lcp": "u1", "llv": "<u8", "bck": "<u4"}
