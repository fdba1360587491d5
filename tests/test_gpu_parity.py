"""GPU: the CUDA path, called through the C-ABI (ctypes -> libgtb200.so), against
  * the outputs of the unmodified reference (tests/golden/reference_vectors.npz),
  * the pinned oracle on seeded inputs,
  * size-independent properties at larger sizes.
Bit-exact: integer/byte work only, no tolerance anywhere."""
import ctypes
import hashlib
import numpy as np
import pytest

import esa_oracle as eo
import synth
from conftest import golden_cases
from genometools_b200 import _lib, encode_symbols
from genometools_b200.suffixerator import Suffixerator, build_esa

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    a = np.frombuffer(a, dtype=np.uint8); b = np.frombuffer(b, dtype=np.uint8)
    if a.shape != b.shape:
        return f"length {a.shape[0]} vs {b.shape[0]}"
    d = np.flatnonzero(a != b)
    return "equal" if d.size == 0 else f"{d.size} bytes differ, first at byte {d[0]}"


def check_against_oracle(sym, K, pl, res, what=""):
    o = eo.esa(sym, K, pl)
    im = eo.file_images(o)
    got = {"suf": res.suf_bytes(), "lcp": res.lcp_bytes(), "llv": res.llv_bytes(), "bck": res.bck_bytes()}
    for ext in ("bck", "suf", "lcp", "llv"):
        assert got[ext] == im[ext], f"{what} .{ext}: {first_diff(got[ext], im[ext])}"
    assert res.longest == o["longest"], what
    assert res.maxbranchdepth == o["maxbranchdepth"], what
    assert res.numoflargelcpvalues == o["numoflargelcpvalues"], what
    assert res.lcptabsum == o["lcptabsum"], what


@pytest.mark.parametrize("case", golden_cases(small_only=True))
def test_cuda_matches_reference_outputs(golden, case):
    m = golden.meta(case)
    sym = golden.symbols(case)
    enc = encode_symbols(sym, m["numofchars"], m["numofsequences"])
    rng = np.random.default_rng(len(case))
    filler = rng.integers(0, 4, size=max(1, int((sym >= 254).sum())), dtype=np.uint8)   # arbitrary filler bases
    res = build_esa(enc, m["prefixlength"], filler=filler if enc.is_dna else None)
    got = {"suf": res.suf_bytes(), "lcp": res.lcp_bytes(), "llv": res.llv_bytes(), "bck": res.bck_bytes()}
    ref = {"suf": golden.get(case, "suf").astype("<u8").tobytes(), "lcp": bytes(golden.get(case, "lcp")),
           "llv": golden.get(case, "llv").astype("<u8").tobytes(), "bck": bytes(golden.get(case, "bck"))}
    for ext in ("bck", "suf", "lcp", "llv"):
        assert got[ext] == ref[ext], f".{ext}: {first_diff(got[ext], ref[ext])}"
        assert hashlib.md5(got[ext]).hexdigest() == str(golden.get(case, "md5_" + ext))
    prj, text = golden.prj(case)
    info = {k: int(prj[k]) for k in ("specialcharacters", "specialranges", "realspecialranges",
                                     "lengthofspecialprefix", "lengthofspecialsuffix", "wildcards",
                                     "wildcardranges", "realwildcardranges", "lengthofwildcardprefix",
                                     "lengthofwildcardsuffix")}   # encseq-derived lines: GtEncseq's job
    assert res.prj_text(info, m["numofsequences"]) == text


def test_cuda_big_golden_md5(golden):
    case = "synth/rand_dna_2M"
    m = golden.meta(case)
    sym = golden.symbols(case)
    res = build_esa(encode_symbols(sym, 4), m["prefixlength"])
    for ext, b in (("suf", res.suf_bytes()), ("lcp", res.lcp_bytes()), ("llv", res.llv_bytes()), ("bck", res.bck_bytes())):
        assert hashlib.md5(b).hexdigest() == str(golden.get(case, "md5_" + ext)), ext


@pytest.mark.parametrize("seed", range(6))
def test_cuda_matches_oracle_random_dna(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 30000))
    sym = synth.random_dna(n, seed, p_n=[0.0, 0.001, 0.05, 0.3, 0.0, 0.9][seed])
    pl = int(rng.integers(1, 7))
    res = build_esa(encode_symbols(sym, 4), pl)
    check_against_oracle(sym, 4, pl, res, f"seed {seed} n {n} pl {pl}")


@pytest.mark.parametrize("name", ["reads", "repeats", "lowcomplex", "protein", "protein_dup", "tiny_alphabet"])
def test_cuda_matches_oracle_shapes(name):
    if name == "reads":
        sym, K, pl = synth.reads(3000, 150, 5, 0.001), 4, 6
    elif name == "repeats":
        sym, K, pl = synth.repeats_dna(300_000, 9, unit=6000, copies=20, exact_len=20000, exact_copies=4), 4, 7
    elif name == "lowcomplex":
        sym, K, pl = synth.low_complexity_dna(100_000, 2), 4, 4
    elif name == "protein":
        sym, K, pl = synth.protein(200_000, 3), 20, 3
    elif name == "protein_dup":
        sym, K, pl = np.concatenate([synth.protein(30_000, 4), [255], synth.protein(30_000, 4)]).astype(np.uint8), 20, 2
    else:
        sym, K, pl = (np.random.default_rng(1).integers(0, 3, 50_000).astype(np.uint8)), 3, 4
    res = build_esa(encode_symbols(sym, K), pl)
    check_against_oracle(sym, K, pl, res, name)


@pytest.mark.parametrize("sym", [[], [254], [255, 254, 255], [0], [3, 3, 3, 3], [0, 254, 0], [2] * 70,
                                 [0, 1] * 40 + [254] + [0, 1] * 40])
def test_cuda_edge_cases(sym):
    s = np.array(sym, dtype=np.uint8)
    res = build_esa(encode_symbols(s, 4), 1)
    check_against_oracle(s, 4, 1, res, str(sym[:8]))


# -parts runs either as a gtb_group (ranges read each other's memory: the multi-GPU product path) or with
# the request/answer rank exchange of genometools_b200/multirange.py (the single-GPU twin of the NCCL protocol)
@pytest.mark.parametrize("protocol", ["group", "exchange"])
@pytest.mark.parametrize("parts", [2, 3, 7])
def test_parts_do_not_change_the_output(monkeypatch, parts, protocol):
    # sfx-partssuf.c: the output is independent of -parts (reference test :64-68)
    monkeypatch.setenv("GTB200_PARTS_PROTOCOL", protocol)
    sym = synth.random_dna(200_000, 77, p_n=0.001)
    enc = encode_symbols(sym, 4)
    one = build_esa(enc, 6)
    many = build_esa(enc, 6, parts=parts)
    assert len(many.stats) == parts
    assert many.numoflargelcpvalues == one.numoflargelcpvalues
    assert one.suf_bytes() == many.suf_bytes()
    assert one.lcp_bytes() == many.lcp_bytes()
    assert one.bck_bytes() == many.bck_bytes()
    assert (one.longest, one.maxbranchdepth, one.lcptabsum) == (many.longest, many.maxbranchdepth, many.lcptabsum)
    check_against_oracle(sym, 4, 6, many, f"parts {parts}")


@pytest.mark.parametrize("name,parts", [("repeats", 2), ("repeats", 5), ("lowcomplex", 3), ("reads_dup", 4),
                                        ("protein_dup", 3)])
@pytest.mark.parametrize("protocol", ["group", "exchange"])
def test_parts_with_ties_across_ranges(monkeypatch, name, parts, protocol):
    """ties whose doubling partner lives in another code range: the rank exchange between
    ranges (the multi-GPU protocol, here with all ranges on one device)"""
    monkeypatch.setenv("GTB200_PARTS_PROTOCOL", protocol)
    if name == "repeats":
        sym, K, pl = synth.repeats_dna(200_000, 19, unit=5000, copies=12, exact_len=9000, exact_copies=4), 4, 5
    elif name == "lowcomplex":
        sym, K, pl = synth.low_complexity_dna(60_000, 8), 4, 3
    elif name == "reads_dup":
        r = synth.reads(500, 120, 3, 0.002)
        sym, K, pl = np.concatenate([r, [255], r, [255], r[:30000]]).astype(np.uint8), 4, 4
    else:
        sym, K, pl = np.concatenate([synth.protein(20_000, 6), [255], synth.protein(20_000, 6)]).astype(np.uint8), 20, 2
    res = build_esa(encode_symbols(sym, K), pl, parts=parts)
    assert len(res.stats) > 1
    assert sum(st["unresolved_after_first_sort"] for st in res.stats) > 0
    check_against_oracle(sym, K, pl, res, f"{name} parts {parts}")


def test_handle_is_reusable_and_deterministic():
    sym = synth.random_dna(150_000, 5, p_n=0.01)
    enc = encode_symbols(sym, 4)
    with Suffixerator(0) as s:
        s.set_sequence(enc)
        a = s.run(6)
        b = s.run(6)
        c = s.run(3)
    assert a.suf_bytes() == b.suf_bytes() == c.suf_bytes()
    assert a.lcp_bytes() == b.lcp_bytes() == c.lcp_bytes()
    assert a.bck_bytes() == b.bck_bytes() != c.bck_bytes()


@pytest.mark.parametrize("n,bits", [(1, (0, 64)), (4096, (0, 64)), (4097, (0, 64)), (1_000_003, (0, 64)),
                                    (300_000, (8, 40)), (200_000, (0, 5))])
def test_onesweep_radix_sort_pairs(n, bits):
    """the stand-alone entry of the sort engine: stable LSD sort on a bit range"""
    lib = _lib.load()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n, dtype=np.uint64)
    if n > 1000:
        keys[::7] = keys[3]                        # many duplicates -> stability matters
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    buf = ctypes.create_string_buffer(256)
    rc = lib.gtb_radixsort_pairs_u64_u32(0, k.ctypes.data, v.ctypes.data, n, bits[0], bits[1], buf, 256)
    assert rc == 0, buf.value
    mask = np.uint64(((1 << (bits[1] - bits[0])) - 1) << bits[0])
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(v, vals[order])
    assert np.array_equal(k, keys[order])


def test_large_input_properties():
    """beyond oracle size: permutation, sortedness of sampled neighbours, lcp consistency"""
    n = 20_000_000
    sym = synth.random_dna(n, 123, p_n=0.0005)
    enc = encode_symbols(sym, 4)
    res = build_esa(enc, 10)
    suf = res.suftab
    assert suf.shape[0] == n + 1 and suf[-1] == n
    seen = np.zeros(n + 1, dtype=np.uint8); seen[suf] = 1
    assert seen.all()                                              # permutation of 0..n
    S = int((sym >= 254).sum())
    assert np.array_equal(suf[n - S:n], np.flatnonzero(sym >= 254).astype(np.uint64))   # special tail
    assert not res.lcptab[n - S:].any()
    run = np.zeros(n + 1, dtype=np.int64)                          # regular run lengths
    idx = np.flatnonzero(sym >= 254)
    nxt = np.full(n + 1, n, dtype=np.int64)
    nxt[idx] = idx
    nxt = np.minimum.accumulate(nxt[::-1])[::-1]
    run = nxt - np.arange(n + 1)
    rng = np.random.default_rng(0)
    for j in rng.integers(1, n - S, size=3000):
        a, b = int(suf[j - 1]), int(suf[j])
        lim = int(min(run[a], run[b]))
        l = 0
        while l < lim and sym[a + l] == sym[b + l]:
            l += 1
        assert min(l, 255) == int(res.lcptab[j]), j
        if l < lim:
            assert sym[a + l] < sym[b + l], j                      # order
        else:
            assert run[a] > run[b] or (run[a] == run[b] and a < b), j
    lb = res.leftborder
    assert lb[0] == 0 and lb[-1] == n - S and np.all(np.diff(lb.astype(np.int64)) >= 0)
    assert res.stats[0]["kernel_launches"] > 0


# every key length (5/6/7/8 radix passes) and every refinement path: text-driven rounds
# only, text rounds followed by prefix doubling (ranks corrected after the text rounds),
# prefix doubling only -- all must give the oracle's bytes
@pytest.mark.parametrize("m,text_rounds", [(17, 0), (17, 1), (17, 8), (21, 2), (25, 0), (29, 2), (29, 0), (9, 1)])
def test_cuda_key_lengths_and_refinement_paths(monkeypatch, m, text_rounds):
    monkeypatch.setenv("GTB200_KEY_SYMBOLS", str(m))
    monkeypatch.setenv("GTB200_TEXT_ROUNDS", str(text_rounds))
    cases = [
        ("repeats", synth.repeats_dna(40_000, 5, unit=1500, copies=5, exact_len=900, exact_copies=3), 5),
        ("reads", synth.reads(300, 60, 7, p_n=0.01), 4),
        ("polyA", np.concatenate([np.zeros(3000, np.uint8), synth.random_dna(2000, 3, p_n=0.01),
                                  np.zeros(2500, np.uint8)]), 3),
    ]
    for name, sym, pl in cases:
        res = build_esa(encode_symbols(sym, 4), pl)
        check_against_oracle(sym, 4, pl, res, f"{name} m={m} text_rounds={text_rounds}")


# the first-level sort without the pass over the tail digit (key lengths that are multiples of four
# symbols: the keys with a tail are sorted apart and appended as a second source of the first pass),
# forced on inputs FULL of specials, and the same formats with the tail pass -- all the oracle's bytes
@pytest.mark.parametrize("m,tail_last,text_rounds", [(16, 1, 0), (16, 1, 2), (20, 1, 0), (24, 1, 1), (28, 1, 0),
                                                     (20, 0, 0), (16, 0, 1), (12, 1, 0), (8, 1, 1)])
def test_cuda_tail_keys_sorted_apart(monkeypatch, m, tail_last, text_rounds):
    monkeypatch.setenv("GTB200_KEY_SYMBOLS", str(m))
    monkeypatch.setenv("GTB200_TAIL_LAST", str(tail_last))
    monkeypatch.setenv("GTB200_TEXT_ROUNDS", str(text_rounds))
    cases = [
        ("repeats", synth.repeats_dna(60_000, 5, unit=1500, copies=6, exact_len=900, exact_copies=3, nruns=5), 5),
        ("reads", synth.reads(500, 60, 7, p_n=0.01), 4),
        ("reads_dup", np.concatenate([synth.reads(200, 40, 2, 0.0), [255], synth.reads(200, 40, 2, 0.0)]).astype(np.uint8), 3),
        ("polyT_N", np.concatenate([np.full(3000, 3, np.uint8), synth.random_dna(2000, 3, p_n=0.02),
                                    np.full(2500, 3, np.uint8), [254], np.full(700, 3, np.uint8)]).astype(np.uint8), 3),
        ("starts_near_special", np.array([1, 2, 254, 0, 3, 3, 3, 255, 3, 3, 1], dtype=np.uint8), 1),
        ("no_specials", synth.random_dna(30_000, 8), 6),
    ]
    for name, sym, pl in cases:
        for parts in (1, 3):
            res = build_esa(encode_symbols(sym, 4), pl, parts=parts)
            check_against_oracle(sym, 4, pl, res, f"{name} m={m} tail_last={tail_last} parts={parts}")


@pytest.mark.parametrize("m,text_rounds", [(8, 0), (8, 2), (10, 1), (12, 2)])
def test_cuda_protein_key_lengths(monkeypatch, m, text_rounds):
    monkeypatch.setenv("GTB200_KEY_SYMBOLS", str(m))
    monkeypatch.setenv("GTB200_TEXT_ROUNDS", str(text_rounds))
    sym = synth.protein(6000, 21, reclen=90, p_x=0.003)
    sym[1000:1400] = sym[3000:3400]            # a 400-residue repeat: lcp >= 255 -> .llv
    for pl in (1, 2):
        res = build_esa(encode_symbols(sym, 20), pl)
        check_against_oracle(sym, 20, pl, res, f"protein m={m} text_rounds={text_rounds} pl={pl}")


@pytest.mark.parametrize("m,tail_last,text_rounds", [(8, 1, 0), (9, 1, 0), (9, 1, 2), (8, 0, 1), (9, 0, 0)])
def test_cuda_protein_tail_keys_sorted_apart(monkeypatch, m, tail_last, text_rounds):
    """byte path: key lengths whose tail field ends on a byte boundary (m = 8; m = 9 with three zero bits
    between the symbols and the tail), the tail keys sorted apart -- and the same lengths with the tail pass"""
    monkeypatch.setenv("GTB200_KEY_SYMBOLS", str(m))
    monkeypatch.setenv("GTB200_TAIL_LAST", str(tail_last))
    monkeypatch.setenv("GTB200_TEXT_ROUNDS", str(text_rounds))
    sym = synth.protein(9000, 23, reclen=70, p_x=0.004)
    sym[1200:1700] = sym[4000:4500]            # a 500-residue repeat: lcp >= 255 -> .llv
    for pl in (1, 2):
        for parts in (1, 3):
            res = build_esa(encode_symbols(sym, 20), pl, parts=parts)
            check_against_oracle(sym, 20, pl, res, f"protein m={m} tail_last={tail_last} pl={pl} parts={parts}")


def test_cuda_partial_counts_and_range_split():
    """count allreduce building blocks: the raw counts of two halves of the text add up to the
    bucket table of the whole, and the device-side range split equals the host mirror of
    gt_suftabparts_new"""
    import ctypes as C
    import torch
    from genometools_b200.multirange import DeviceArray
    from genometools_b200.sharding import suftab_parts
    lib = _lib.load()
    sym = synth.reads(400, 70, 9, p_n=0.02)
    enc = encode_symbols(sym, 4)
    n, pl = enc.totallength, 5
    words, ranges = enc.twobitencoding()
    buf = C.create_string_buffer(512)
    hs = [lib.gtb_esa_new(0, buf, 512) for _ in range(3)]
    for h in hs:
        assert lib.gtb_esa_set_input_2bit(h, _lib.ptr(words), words.shape[0], n, _lib.ptr(ranges) if ranges.shape[0] else None,
                                          ranges.shape[0]) == 0
    nall, nsp, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.gtb_bck_sizes(4, pl, C.byref(nall), C.byref(nsp), C.byref(nd))
    sizes = (nall.value + 1, nsp.value, nd.value)

    def tables(h):
        p = [C.c_void_p() for _ in range(3)]
        assert lib.gtb_esa_dev_bcktab(h, *[C.byref(x) for x in p]) == 0
        return [torch.as_tensor(DeviceArray(x.value, c, "<i4"), device="cuda") for x, c in zip(p, sizes)]

    cut = n // 3
    assert lib.gtb_esa_count_partial(hs[0], pl, 0, cut) == 0
    assert lib.gtb_esa_count_partial(hs[1], pl, cut, n) == 0
    for a, b in zip(tables(hs[0]), tables(hs[1])):
        a += b                                       # what the all-reduce does
    torch.cuda.synchronize()
    assert lib.gtb_esa_count_finish(hs[0]) == 0
    assert lib.gtb_esa_count(hs[2], pl) == 0
    got, ref = [], []
    for h, dst in ((hs[0], got), (hs[2], ref)):
        lb = np.empty(sizes[0], np.uint32); cs = np.empty(sizes[1], np.uint32); di = np.empty(max(sizes[2], 1), np.uint32)
        assert lib.gtb_esa_copy_bcktab(h, _lib.ptr(lb), _lib.ptr(cs), _lib.ptr(di)) == 0
        dst.extend([lb, cs, di[:sizes[2]]])
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    for parts in (1, 2, 3, 5, 8, 64):
        out = (C.c_uint64 * (4 * parts))()
        k = C.c_uint()
        assert lib.gtb_esa_split_ranges(hs[0], parts, out, C.byref(k)) == 0
        mine = [tuple(int(out[4 * p + q]) for q in range(4)) for p in range(k.value)]
        assert mine == suftab_parts(ref[0], parts), parts
    for h in hs:
        lib.gtb_esa_delete(h)


def test_cuda_slice_partition_and_pair_exchange():
    """the sharded text scan on one GPU: two "ranks" partition their halves of the text by
    owning code range, the groups are concatenated per owner in rank order (what the
    all-to-all delivers) and each range sorts its pairs -- same bytes as the oracle"""
    import ctypes as C
    import torch
    from genometools_b200.multirange import GpuRangeWorker, run_ranges_local, range_first_keys
    from genometools_b200._lib import GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, GTB_REUSE_COUNTS, ptr
    from genometools_b200.sharding import suftab_parts
    lib = _lib.load()
    for sym, pl, R in ((synth.repeats_dna(50_000, 8, unit=1500, copies=5, exact_len=900, exact_copies=3), 5, 2),
                       (synth.reads(500, 70, 3, p_n=0.02), 4, 3)):
        enc = encode_symbols(sym, 4)
        n = enc.totallength
        words, ranges = enc.twobitencoding()
        buf = C.create_string_buffer(512)
        h0 = lib.gtb_esa_new(0, buf, 512)
        assert lib.gtb_esa_set_input_2bit(h0, ptr(words), words.shape[0], n, ptr(ranges) if ranges.shape[0] else None,
                                          ranges.shape[0]) == 0
        assert lib.gtb_esa_count(h0, pl) == 0
        nall, nsp, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(4, pl, C.byref(nall), C.byref(nsp), C.byref(nd))
        lb = np.empty(nall.value + 1, np.uint32)
        assert lib.gtb_esa_copy_bcktab(h0, ptr(lb), None, None) == 0
        parts = suftab_parts(lb, R)
        assert len(parts) == R
        fk = range_first_keys(4, pl, parts)
        groups = [[] for _ in range(R)]          # groups[owner] = list over slices
        for r in range(R):
            lo, hi = n * r // R, n * (r + 1) // R
            sk = torch.empty(max(hi - lo, 1), dtype=torch.int64, device="cuda")
            sp = torch.empty(max(hi - lo, 1), dtype=torch.int32, device="cuda")
            counts = np.zeros(R, np.uint64)
            assert lib.gtb_esa_slice_partition(h0, pl, lo, hi, ptr(fk), R, sk.data_ptr(), sp.data_ptr(), max(hi - lo, 1),
                                               ptr(counts)) == 0, lib.gtb_esa_error(h0)
            off = 0
            for g in range(R):
                c = int(counts[g])
                groups[g].append((sk[off:off + c].clone(), sp[off:off + c].clone()))
                off += c
        flags = GTB_WANT_SUF | GTB_WANT_LCP | GTB_WANT_BCK | GTB_REUSE_COUNTS
        hs, workers, begins, keep = [], [], [], []
        for g, (mn, mx, off, _w) in enumerate(parts):
            h = lib.gtb_esa_new(0, buf, 512)
            hs.append(h)
            assert lib.gtb_esa_share_input(h, h0) == 0
            assert lib.gtb_esa_count(h, pl) == 0
            assert lib.gtb_esa_set_code_range(h, mn, mx, off, 1 if g == R - 1 else 0) == 0
            k = torch.cat([x[0] for x in groups[g]]); p = torch.cat([x[1] for x in groups[g]])
            keep.append((k, p))
            workers.append(GpuRangeWorker(h, pl, flags, 0))
            begins.append(lambda h=h, k=k, p=p: _ck0(lib, h, lib.gtb_esa_sort_begin_pairs(h, pl, flags, k.data_ptr(),
                                                                                         p.data_ptr(), k.numel())))
        run_ranges_local(workers, fk, True, begins)
        suf, lcp = [], []
        for h in hs:
            e = lib.gtb_esa_num_entries(h)
            a = np.empty(e, np.uint64); b = np.empty(e, np.uint8)
            assert lib.gtb_esa_copy_suftab_u64(h, ptr(a), 0, e) == 0
            assert lib.gtb_esa_copy_lcptab(h, ptr(b), 0, e) == 0
            suf.append(a); lcp.append(b)
        o = eo.esa(sym, 4, pl)
        im = eo.file_images(o)
        assert np.concatenate(suf).astype("<u8").tobytes() == im["suf"]
        assert np.concatenate(lcp).tobytes() == im["lcp"]
        for h in hs + [h0]:
            lib.gtb_esa_delete(h)


def _ck0(lib, h, rc):
    assert rc == 0, lib.gtb_esa_error(h).decode()


@pytest.mark.parametrize("exchange", [False, True, "positions"])
def test_cuda_coarse_ranges_and_merged_bucket_table(exchange):
    """the multi-GPU flow on one GPU: coarse counts -> ranges of whole coarse buckets -> every
    range sorts (from the text, or from exchanged pairs) and fills the bucket-table entries
    of its own codes from its sorted keys; the summed tables and the concatenated suffix /
    lcp tables equal the oracle's"""
    import ctypes as C
    import torch
    from genometools_b200.multirange import GpuRangeWorker, run_ranges_local, range_first_keys
    from genometools_b200._lib import GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, GTB_REUSE_COUNTS, ptr
    lib = _lib.load()
    for sym, pl, R in ((synth.repeats_dna(60_000, 18, unit=1500, copies=5, exact_len=900, exact_copies=3), 8, 3),
                       (synth.reads(600, 70, 13, p_n=0.02), 7, 2), (synth.random_dna(30_000, 4, 0.01), 3, 4),
                       (synth.low_complexity_dna(20_000, 5), 3, 4),      # ties whose partners end middle ranges
                       (synth.repeats_dna(40_000, 2, unit=900, copies=12, exact_len=2500, exact_copies=4), 2, 5)):
        enc = encode_symbols(sym, 4)
        n = enc.totallength
        words, ranges = enc.twobitencoding()
        buf = C.create_string_buffer(512)
        h0 = lib.gtb_esa_new(0, buf, 512)
        assert lib.gtb_esa_set_input_2bit(h0, ptr(words), words.shape[0], n, ptr(ranges) if ranges.shape[0] else None,
                                          ranges.shape[0]) == 0
        pd, nc = C.c_void_p(), C.c_uint64()
        assert lib.gtb_esa_coarse_partial(h0, pl, 0, n, C.byref(pd), C.byref(nc)) == 0
        out4 = (C.c_uint64 * (4 * R))(); k = C.c_uint()
        assert lib.gtb_esa_coarse_split(h0, R, out4, C.byref(k)) == 0
        parts = [tuple(int(out4[4 * p + q]) for q in range(4)) for p in range(k.value)]
        assert parts[0][0] == 0 and parts[-1][1] == 4 ** pl - 1 and sum(p[3] for p in parts) == int((sym < 4).sum())
        fk = range_first_keys(4, pl, parts)
        flags = GTB_WANT_SUF | GTB_WANT_LCP | GTB_WANT_BCK | GTB_REUSE_COUNTS
        hs, workers, begins, keep = [], [], [], []
        groups = [[] for _ in parts]
        if exchange:
            for r in range(2):                        # two slices of the text
                lo, hi = n * r // 2, n * (r + 1) // 2
                sk = torch.empty(max(hi - lo, 1), dtype=torch.int64, device="cuda")
                sp = torch.empty(max(hi - lo, 1), dtype=torch.int32, device="cuda")
                counts = np.zeros(len(parts), np.uint64)
                assert lib.gtb_esa_slice_partition(h0, pl, lo, hi, ptr(fk), len(parts),
                                                   None if exchange == "positions" else sk.data_ptr(), sp.data_ptr(),
                                                   max(hi - lo, 1), ptr(counts)) == 0
                off = 0
                for g in range(len(parts)):
                    c = int(counts[g]); groups[g].append((sk[off:off + c].clone(), sp[off:off + c].clone())); off += c
        for g, (mn, mx, off, width) in enumerate(parts):
            h = lib.gtb_esa_new(0, buf, 512)
            hs.append(h)
            assert lib.gtb_esa_share_input(h, h0) == 0
            assert lib.gtb_esa_set_code_range_known(h, mn, mx, off, width, 1 if g == len(parts) - 1 else 0) == 0
            workers.append(GpuRangeWorker(h, pl, flags, 0))
            if exchange:
                kk = torch.cat([x[0] for x in groups[g]]); pp = torch.cat([x[1] for x in groups[g]])
                keep.append((kk, pp))
                if exchange == "positions":
                    begins.append(lambda h=h, pp=pp: _ck0(lib, h, lib.gtb_esa_sort_begin_positions(
                        h, pl, flags, pp.data_ptr(), pp.numel())))
                else:
                    begins.append(lambda h=h, kk=kk, pp=pp: _ck0(lib, h, lib.gtb_esa_sort_begin_pairs(
                        h, pl, flags, kk.data_ptr(), pp.data_ptr(), kk.numel())))
            else:
                begins.append(None)
        run_ranges_local(workers, fk, True, begins)
        nall, nsp, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(4, pl, C.byref(nall), C.byref(nsp), C.byref(nd))
        tabs = [np.zeros(nall.value + 1, np.uint64), np.zeros(nsp.value, np.uint64), np.zeros(max(nd.value, 1), np.uint64)]
        suf, lcp = [], []
        for h in hs:
            e = lib.gtb_esa_num_entries(h)
            a = np.empty(e, np.uint64); b = np.empty(e, np.uint8)
            assert lib.gtb_esa_copy_suftab_u64(h, ptr(a), 0, e) == 0
            assert lib.gtb_esa_copy_lcptab(h, ptr(b), 0, e) == 0
            suf.append(a); lcp.append(b)
            t = [np.empty(nall.value + 1, np.uint32), np.empty(nsp.value, np.uint32), np.empty(max(nd.value, 1), np.uint32)]
            assert lib.gtb_esa_copy_bcktab(h, ptr(t[0]), ptr(t[1]), ptr(t[2]) if nd.value else None) == 0
            for x, y in zip(tabs, t):
                x += y                                # what the all-reduce over the ranks does
        o = eo.esa(sym, 4, pl)
        im = eo.file_images(o)
        assert np.concatenate(suf).astype("<u8").tobytes() == im["suf"]
        assert np.concatenate(lcp).tobytes() == im["lcp"]
        assert np.array_equal(tabs[0], o["leftborder"].astype(np.uint64))
        assert np.array_equal(tabs[1], o["countspecialcodes"].astype(np.uint64))
        assert np.array_equal(tabs[2][:nd.value], o["distpfxidx"].astype(np.uint64))
        for h in hs + [h0]:
            lib.gtb_esa_delete(h)


@pytest.mark.parametrize("mode", ["", "wide", "both"])
@pytest.mark.parametrize("stage_kb", ["64", "8192"])
def test_result_copy_paths_agree(mode, stage_kb, monkeypatch):
    """gtb_esa_copy_suftab_u64 into PINNED memory: staged uint32 + host widening (default), widened
    on the device (GTB200_SUF_COPY=wide), both dealt from the two ends (=both); staging chunks
    smaller and larger than the table.  All must deliver the table the pageable path delivers."""
    import torch
    if mode:
        monkeypatch.setenv("GTB200_SUF_COPY", mode)
    monkeypatch.setenv("GTB200_STAGE_KB", stage_kb)
    monkeypatch.setenv("GTB200_HOST_THREADS", "5")
    lib = _lib.load()
    n = 3_000_001                                    # 12 MB of uint32: 184 chunks of 64 KiB, 2 of 8 MiB
    sym = synth.random_dna(n, 77, p_n=0.001)
    enc = encode_symbols(sym, 4)
    with Suffixerator(0) as s:
        s.set_sequence(enc)
        res = s.run(8, want_bck=True)                # the library's pageable path (numpy buffers)
        e = int(lib.gtb_esa_num_entries(s.h))
        for first, count in ((0, e), (12345, e - 20000), (e - 1, 1), (7, 0)):
            pinned = torch.empty(max(count, 1), dtype=torch.int64, pin_memory=True)
            pinned.fill_(-1)
            assert lib.gtb_esa_copy_suftab_u64(s.h, pinned.data_ptr(), first, count) == 0, lib.gtb_esa_error(s.h)
            got = pinned.numpy().view(np.uint64)[:count]
            assert np.array_equal(got, res.suftab[first:first + count]), (mode, first, count)
            lp = torch.empty(max(count, 1), dtype=torch.uint8, pin_memory=True)
            assert lib.gtb_esa_copy_lcptab(s.h, lp.data_ptr(), first, count) == 0
            assert np.array_equal(lp.numpy()[:count], res.lcptab[first:first + count])
            if first == 0 and count == e:                 # every table in one call, small ones on a second stream
                pinned.fill_(-1); lp.fill_(7)
                lbp = torch.empty(res.leftborder.shape[0], dtype=torch.int32, pin_memory=True)
                cs = np.empty_like(res.countspecialcodes)                 # (pageable: served afterwards)
                assert lib.gtb_esa_copy_results(s.h, pinned.data_ptr(), lp.data_ptr(), None, lbp.data_ptr(),
                                                cs.ctypes.data, None) == 0, lib.gtb_esa_error(s.h)
                assert np.array_equal(pinned.numpy().view(np.uint64)[:count], res.suftab)
                assert np.array_equal(lp.numpy()[:count], res.lcptab)
                assert np.array_equal(lbp.numpy().view(np.uint32), res.leftborder)
                assert np.array_equal(cs, res.countspecialcodes)
            pinned.fill_(-1); lp.fill_(7)                 # both tables in one call (lcp on a second stream)
            assert lib.gtb_esa_copy_tables(s.h, pinned.data_ptr(), lp.data_ptr(), first, count) == 0
            assert np.array_equal(pinned.numpy().view(np.uint64)[:count], res.suftab[first:first + count])
            assert np.array_equal(lp.numpy()[:count], res.lcptab[first:first + count])
