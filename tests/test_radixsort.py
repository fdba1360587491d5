"""The record sorts of src/core/radix_sort.h ("next" row 4) on the onesweep
engine: gtb_radixsort_u64 / _u64pair / _u64keypair stand in for gt_radixsort_inplace_ulong /
_GtUwordPair / _Gtuint64keyPair (/root/reference/src/core/radix_sort.h:91,107,125).
CPU: numpy's sort (the oracle) against md5 sums of the unmodified reference's outputs
(tests/golden/radixsort_vectors.npz, made by tests/golden/make_golden_radixsort.py).
GPU: the CUDA path against both."""
import ctypes
import hashlib
import os
import numpy as np
import pytest

from conftest import ROOT
from make_golden_radixsort import CASES, FLBA_CASES, make_flba, make_input
from genometools_b200 import _lib

Z = np.load(os.path.join(ROOT, "tests", "golden", "radixsort_vectors.npz"), allow_pickle=False)


def oracle_sort(kind, a):
    if kind == "ulong":
        return np.sort(a, axis=0)
    if kind == "keypair":
        return a[np.lexsort((a[:, 1], a[:, 0]))]
    return a[np.argsort(a[:, 0], kind="stable")]          # key = first component


def md5(x):
    return hashlib.md5(np.ascontiguousarray(x).tobytes()).hexdigest()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_sort_matches_reference(name):
    kind, n, seed, distinct = CASES[name]
    a = make_input(kind, n, seed, distinct)
    b = oracle_sort(kind, a)
    assert md5(b[:, 0]) == str(Z[name + "/md5_keys"])
    if kind != "ulongpair":            # (the reference leaves the order of equal keys of a pair unspecified)
        assert md5(b) == str(Z[name + "/md5"])


def gpu_sort(kind, a):
    lib = _lib.load()
    fn = {"ulong": lib.gtb_radixsort_u64, "ulongpair": lib.gtb_radixsort_u64pair,
          "keypair": lib.gtb_radixsort_u64keypair}[kind]
    b = np.ascontiguousarray(a.copy())
    buf = ctypes.create_string_buffer(256)
    assert fn(0, b.ctypes.data, b.shape[0], buf, 256) == 0, buf.value
    return b


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_sort_matches_reference(name):
    kind, n, seed, distinct = CASES[name]
    a = make_input(kind, n, seed, distinct)
    b = gpu_sort(kind, a)
    assert np.array_equal(b, oracle_sort(kind, a))
    assert md5(b[:, 0]) == str(Z[name + "/md5_keys"])
    if kind != "ulongpair":
        assert md5(b) == str(Z[name + "/md5"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ulong", "ulongpair", "keypair"])
@pytest.mark.parametrize("n", [0, 1, 2, 6143, 6144, 6145, 1_000_003])
def test_cuda_sort_sizes(kind, n):
    rng = np.random.default_rng(n + len(kind))
    a = rng.integers(0, 2 ** 63, size=(n, 1 if kind == "ulong" else 2), dtype=np.uint64)
    if n > 100:
        a[::3, 0] = a[1, 0]                                  # ties: stability / second key matter
        a[::7] = a[2]
    assert np.array_equal(gpu_sort(kind, a), oracle_sort(kind, a))


RADIX_BIN = os.path.join(ROOT, "host", "_build", "gtref_b200_radixsort")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RADIX_BIN), reason="host/_build/gtref_b200_radixsort not built")
@pytest.mark.parametrize("name", list(CASES))
def test_reference_entry_points_on_the_gpu(tmp_path, name):
    """host/gt_radix_sort_b200.c linked in front of src/core/radix_sort.o: gt_radixsort_inplace_ulong /
    _GtUwordPair / _Gtuint64keyPair called by the reference's test driver, sorted by libgtb200"""
    import subprocess
    kind, n, seed, distinct = CASES[name]
    a = make_input(kind, n, seed, distinct)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.ascontiguousarray(a).tofile(fin)
    subprocess.check_call([RADIX_BIN, "radixsort", kind, fin, fout], stdout=subprocess.DEVNULL)
    b = np.fromfile(fout, dtype=np.uint64).reshape(a.shape)
    assert np.array_equal(b, oracle_sort(kind, a))
    assert md5(b[:, 0] if b.ndim > 1 else b) == str(Z[name + "/md5_keys"])


def flba_oracle(a):
    return a[np.lexsort(a.T[::-1])] if a.shape[0] > 1 else a          # memcmp order


@pytest.mark.parametrize("name", list(FLBA_CASES))
def test_flba_oracle_matches_reference(name):
    """gt_radixsort_inplace_flba (radix_sort.h:138): records of unitsize bytes, most significant byte first"""
    a = make_flba(*FLBA_CASES[name])
    assert md5(flba_oracle(a)) == str(Z[name + "/md5"])


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RADIX_BIN), reason="host/_build/gtref_b200_radixsort not built")
@pytest.mark.parametrize("name", list(FLBA_CASES))
def test_flba_entry_point_on_the_gpu(tmp_path, name):
    """gt_radixsort_inplace_flba of host/gt_radix_sort_b200.c under the reference's test driver: the records
    travel as big-endian 64-bit keys (up to 8 bytes) or key pairs (up to 16) through libgtb200"""
    import subprocess
    unitsize = FLBA_CASES[name][0]
    a = make_flba(*FLBA_CASES[name])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    a.tofile(fin)
    subprocess.check_call([RADIX_BIN, "radixsort", "flba%d" % unitsize, fin, fout], stdout=subprocess.DEVNULL)
    b = np.fromfile(fout, dtype=np.uint8).reshape(a.shape)
    assert np.array_equal(b, flba_oracle(a))
    assert md5(b) == str(Z[name + "/md5"])


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RADIX_BIN), reason="host/_build/gtref_b200_radixsort not built")
def test_flba_longer_than_16_bytes_is_refused(tmp_path):
    import subprocess
    a = make_flba(17, 100, 1, 0)
    fin = str(tmp_path / "in.bin")
    a.tofile(fin)
    r = subprocess.run([RADIX_BIN, "radixsort", "flba17", fin, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "at most 16" in r.stderr
