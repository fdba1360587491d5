"""CPU, world_size 2, gloo: the lock-step protocol of genometools_b200/multirange.py
(the N > 1 path of bench.py and of the multi-GPU suffixerator) with a numpy stand-in for
the per-range worker.  Checks that the concatenated ranges equal the oracle's suffix
table, that ties crossing the range border are resolved through the rank exchange, and
that the seam lcp is the oracle's."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, sym, K, pl, parts, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_range_worker import CpuRangeWorker
        from genometools_b200.multirange import run_range_distributed
        first_keys = np.array([p[0] << (64 - 2 * pl) for p in parts], dtype=np.uint64)
        nxt = int(first_keys[rank + 1]) if rank + 1 < world else None
        w = CpuRangeWorker(sym, 0 if rank == 0 else int(first_keys[rank]), nxt, parts[rank][2])
        rounds = run_range_distributed(w, first_keys, dist, torch.device("cpu"))
        q.put((rank, list(w.sa), w.lcp0, rounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["repeat_across_border", "random"])
def test_lockstep_protocol_world2(case):
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import esa_oracle as eo
    from genometools_b200.sharding import suftab_parts
    rng = np.random.default_rng(5)
    if case == "random":
        sym = rng.integers(0, 4, size=600, dtype=np.uint8)
        sym[[50, 300, 301]] = 254
    else:
        unit = rng.integers(0, 4, size=120, dtype=np.uint8)        # copies -> ties deeper than 29 everywhere
        sym = np.concatenate([unit, [254], unit, rng.integers(0, 4, size=40, dtype=np.uint8), unit]).astype(np.uint8)
    K, pl, world = 4, 2, 2
    o = eo.esa(sym, K, pl)
    parts = suftab_parts(o["leftborder"], world)
    assert len(parts) == world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sym, K, pl, parts, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nreg = int((sym < 254).sum())
    sa = out[0][1] + out[1][1]
    assert sa == [int(x) for x in o["suf"][:nreg]]
    assert out[1][2] == int(o["lcp"][len(out[0][1])])                # seam lcp
    assert out[0][3] == out[1][3]                                      # same number of rounds on both ranks
    if case == "repeat_across_border":
        assert out[0][3] >= 2


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from genometools_b200.multirange import DistAllgather
        ag = DistAllgather(dist, torch.device("cpu"))
        ok = True
        # the block sizes gtb_esa_run_sharded uses: a status word, the tied count, the 16.5 KB coarse table
        for nbytes in (4, 136, 16512, 40, 70000, 8):
            mine = (np.arange(nbytes, dtype=np.uint32) * 7 + rank * 131 + nbytes).astype(np.uint8)
            out = np.zeros(nbytes * world, dtype=np.uint8)
            rc = ag.fn(None, mine.ctypes.data, nbytes, out.ctypes.data)
            ok = ok and rc == 0
            for r in range(world):
                exp = (np.arange(nbytes, dtype=np.uint32) * 7 + r * 131 + nbytes).astype(np.uint8)
                ok = ok and bool((out[r * nbytes:(r + 1) * nbytes] == exp).all())
        q.put((rank, ok, ag.calls))
    finally:
        dist.destroy_process_group()


def test_sharded_entry_allgather_over_gloo_world2():
    """the one collective gtb_esa_run_sharded needs from its caller (multirange.DistAllgather), driven
    through the C callback type exactly as the library calls it, world_size 2 over gloo"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, True, 6), (1, True, 6)]
