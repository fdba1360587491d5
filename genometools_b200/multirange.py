"""Prefix doubling across bucket-code ranges: the lock-step protocol that lets several
ranges (one per GPU, or several on one GPU for -parts) sort their buckets independently
and still refine ties by rank.

A range owns the ranks (inverse suffix array entries) of its own suffixes.  In doubling
round r (h = m * 2^r) a tied suffix p needs rank(p + h); positions owned by another range
are sent to the owner, looked up there, and the ranks are sent back:

    all-to-all(positions)  ->  rank lookup on the owner  ->  all-to-all(ranks)

The only other exchanges are a max-reduction of the number of tied suffixes (loop
condition) and the boundary keys for the seam lcp.  The reference has no counterpart:
its sorters compare text (src/core/encseq.c:6719) inside one address space; the
partitioning itself mirrors gt_suftabparts_new (src/match/sfx-partssuf.c:172-347).

`RangeWorker` is the interface; GpuRangeWorker binds it to libgtb200.so.  The
orchestration is backend-agnostic (torch.distributed with NCCL on GPUs; the tests drive
it with gloo and a CPU stand-in worker).
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import GtbError, GtbStats, ptr, ALLGATHER_FN


class RangeWorker:
    """what the lock-step driver needs from one code range"""

    def sort_begin(self): raise NotImplementedError
    def unresolved(self): raise NotImplementedError
    def ensure_ranks(self): raise NotImplementedError
    def round_prepare(self, first_keys, my_range): raise NotImplementedError   # -> (positions tensor, counts list)
    def rank_lookup(self, positions): raise NotImplementedError                # -> ranks tensor
    def round_finish(self, answers): raise NotImplementedError
    def sort_end(self): raise NotImplementedError
    def boundary_keys(self): raise NotImplementedError                         # -> (nonempty, first, last)
    def fix_seam(self, prev_last_key): raise NotImplementedError


class GpuRangeWorker(RangeWorker):
    """one code range on one CUDA device, driven through the staged C-ABI"""

    def __init__(self, handle, prefixlength, flags, device):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.h = handle
        self.pl = prefixlength
        self.flags = flags
        self.device = torch.device("cuda", device)
        self.send = None
        self.sent = 0

    def _ck(self, rc):
        if rc != 0:
            raise GtbError(self.lib.gtb_esa_error(self.h).decode())

    def sort_begin(self):
        self._ck(self.lib.gtb_esa_sort_begin(self.h, self.pl, self.flags))

    def unresolved(self):
        return int(self.lib.gtb_esa_unresolved(self.h))

    def ensure_ranks(self):
        self._ck(self.lib.gtb_esa_ensure_ranks(self.h))

    def round_prepare(self, first_keys, my_range):
        torch = self.torch
        M = self.unresolved()
        if self.send is None or self.send.numel() < max(M, 1):
            self.send = torch.empty(max(M, 1), dtype=torch.int32, device=self.device)
        fk = np.ascontiguousarray(first_keys, dtype=np.uint64)
        counts = np.zeros(fk.shape[0], dtype=np.uint64)
        self._ck(self.lib.gtb_esa_round_prepare(self.h, ptr(fk), fk.shape[0], my_range, self.send.data_ptr(),
                                                self.send.numel(), ptr(counts)))
        self.sent = int(counts.sum())
        return self.send[: self.sent], [int(c) for c in counts]

    def rank_lookup(self, positions):
        torch = self.torch
        out = torch.empty(max(positions.numel(), 1), dtype=torch.int32, device=self.device)[: positions.numel()]
        if positions.numel():
            assert positions.is_contiguous()
            self._ck(self.lib.gtb_esa_rank_lookup(self.h, positions.data_ptr(), positions.numel(), out.data_ptr()))
        return out

    def round_finish(self, answers):
        assert answers.numel() == self.sent
        if answers.numel():
            assert answers.is_contiguous()
        self._ck(self.lib.gtb_esa_round_finish(self.h, answers.data_ptr() if answers.numel() else None))

    def sort_end(self):
        self._ck(self.lib.gtb_esa_sort_end(self.h))

    def stats(self):
        st = GtbStats()
        self._ck(self.lib.gtb_esa_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def boundary_keys(self):
        fk, lk = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.gtb_esa_boundary_keys(self.h, C.byref(fk), C.byref(lk)))
        return self.stats()["nonspecials"] > 0, fk.value, lk.value

    def fix_seam(self, prev_last_key):
        self._ck(self.lib.gtb_esa_fix_seam(self.h, prev_last_key))


def range_first_keys(numofchars, prefixlength, parts):
    lib = _lib.load()
    return np.array([lib.gtb_code_first_key(numofchars, prefixlength, p[0]) for p in parts], dtype=np.uint64)


class DeviceArray:
    """a device pointer of the library seen as a CUDA array (torch.as_tensor accepts it)"""

    def __init__(self, p, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (p, False), "version": 2}


def count_allreduce_and_split(lib, handle, numofchars, prefixlength, totallength, dist, device):
    """the bucket table of a text replicated on every rank: rank r counts the k-mers of its
    1/world slice of the text positions, one all-reduce (NCCL over NVLink) per table sums the
    raw counts in place, every rank finishes with the same partial sums and cuts the same
    `world` code ranges (gt_suftabparts_new).  Returns [(mincode, maxcode, sa_offset, width)]."""
    import torch
    world, me = dist.get_world_size(), dist.get_rank()

    def ck(rc):
        if rc != 0:
            raise GtbError(lib.gtb_esa_error(handle).decode())

    lo, hi = totallength * me // world, totallength * (me + 1) // world
    ck(lib.gtb_esa_count_partial(handle, prefixlength, lo, hi))
    nall, nspecial, ndist = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.gtb_bck_sizes(numofchars, prefixlength, C.byref(nall), C.byref(nspecial), C.byref(ndist))
    plb, pcs, pdi = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ck(lib.gtb_esa_dev_bcktab(handle, C.byref(plb), C.byref(pcs), C.byref(pdi)))
    for p, cnt in ((plb, nall.value + 1), (pcs, nspecial.value), (pdi, ndist.value)):
        if cnt and p.value:
            # uint32 counters summed as int32: the same bits
            t = torch.as_tensor(DeviceArray(p.value, cnt, "<i4"), device=device)
            dist.all_reduce(t)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    ck(lib.gtb_esa_count_finish(handle))
    out = (C.c_uint64 * (4 * world))()
    npart = C.c_uint()
    ck(lib.gtb_esa_split_ranges(handle, world, out, C.byref(npart)))
    return [(int(out[4 * p]), int(out[4 * p + 1]), int(out[4 * p + 2]), int(out[4 * p + 3]))
            for p in range(npart.value)]


def coarse_allreduce_and_split(lib, handle, prefixlength, totallength, dist, device):
    """code ranges without a fine-grained counting pass: rank r counts the first few symbols
    of the filled keys (<= 4096 coarse codes, in shared memory) over its 1/world slice of the
    text, one all-reduce sums the tiny table, every rank cuts the same `world` ranges of whole
    coarse buckets.  Returns [(mincode, maxcode, sa_offset, width)] in fine-grained codes."""
    import torch
    world, me = dist.get_world_size(), dist.get_rank()

    def ck(rc):
        if rc != 0:
            raise GtbError(lib.gtb_esa_error(handle).decode())

    lo, hi = totallength * me // world, totallength * (me + 1) // world
    pd, nc = C.c_void_p(), C.c_uint64()
    ck(lib.gtb_esa_coarse_partial(handle, prefixlength, lo, hi, C.byref(pd), C.byref(nc)))
    t = torch.as_tensor(DeviceArray(pd.value, nc.value, "<i4"), device=device)
    dist.all_reduce(t)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    out = (C.c_uint64 * (4 * world))()
    npart = C.c_uint()
    ck(lib.gtb_esa_coarse_split(handle, world, out, C.byref(npart)))
    return [(int(out[4 * p]), int(out[4 * p + 1]), int(out[4 * p + 2]), int(out[4 * p + 3]))
            for p in range(npart.value)]


def allreduce_bcktab(lib, handle, numofchars, prefixlength, dist, device):
    """every rank's run filled the bucket-table entries of its own codes only: sum the three
    tables over the ranks in place (afterwards every rank holds the whole table)"""
    import torch
    nall, nspecial, ndist = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.gtb_bck_sizes(numofchars, prefixlength, C.byref(nall), C.byref(nspecial), C.byref(ndist))
    plb, pcs, pdi = C.c_void_p(), C.c_void_p(), C.c_void_p()
    if lib.gtb_esa_dev_bcktab(handle, C.byref(plb), C.byref(pcs), C.byref(pdi)) != 0:
        raise GtbError("no bucket table on this handle")
    for p, cnt in ((plb, nall.value + 1), (pcs, nspecial.value), (pdi, ndist.value)):
        if cnt and p.value:
            dist.all_reduce(torch.as_tensor(DeviceArray(p.value, cnt, "<i4"), device=device))
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()


# ---------------------------------------------------------------- all ranges in one process
def run_ranges_local(workers, first_keys, want_lcp=True, begins=None):
    """lock-step over ranges that live in this process (same GPU): positions and ranks
    are read in place, nothing is copied.  begins[r] (optional) replaces worker r's
    sort_begin (a range whose pairs were produced by slice partitions)."""
    R = len(workers)
    for r, w in enumerate(workers):
        if begins is not None and begins[r] is not None:
            begins[r]()
        else:
            w.sort_begin()
    if any(w.unresolved() > 0 for w in workers):
        for w in workers:
            w.ensure_ranks()
        rounds = 0
        while any(w.unresolved() > 0 for w in workers):
            prepared = [w.round_prepare(first_keys, r) for r, w in enumerate(workers)]
            answers = []
            for r, (send, counts) in enumerate(prepared):
                parts, off = [], 0
                for o in range(R):
                    if counts[o]:
                        parts.append(workers[o].rank_lookup(send[off: off + counts[o]].contiguous()))
                        off += counts[o]
                if parts:
                    ans = parts[0] if len(parts) == 1 else parts[0].new_empty(off)
                    if len(parts) > 1:
                        o2 = 0
                        for ptn in parts:
                            ans[o2: o2 + ptn.numel()] = ptn
                            o2 += ptn.numel()
                else:
                    ans = send[:0]
                answers.append(ans)
            for w, ans in zip(workers, answers):
                w.round_finish(ans)
            rounds += 1
            if rounds > 64:
                raise GtbError("prefix doubling across ranges did not converge")
    for w in workers:
        w.sort_end()
    if want_lcp:
        prev = None
        for w in workers:
            nonempty, first, last = w.boundary_keys()
            if nonempty:
                if prev is not None:
                    w.fix_seam(prev)
                prev = last


# ---------------------------------------------------------------- one range per process
class PairExchange:
    """Sharded text scan: this rank generates the (filled key, position) pairs of its 1/world
    slice of the text positions, grouped by owning code range (gtb_esa_slice_partition: one
    onesweep pass whose digit is the owner), all_to_all moves every group to its owner --
    the owner receives its pairs in text order (slices in rank order) -- and the owner's sort
    starts from those pairs (gtb_esa_sort_begin_pairs): no rank scans the whole text."""

    def __init__(self, lib, handle, prefixlength, flags, totallength, dist, device, positions_only=True):
        import torch
        self.torch, self.lib, self.h, self.pl, self.flags = torch, lib, handle, prefixlength, flags
        self.n, self.dist, self.device = totallength, dist, device
        self.positions_only = positions_only     # 4 instead of 12 bytes per suffix over the links;
        self.buf = {}                            # the owner regenerates the keys

    def _tensor(self, name, count, dtype):
        t = self.buf.get(name)
        if t is None or t.numel() < count:
            t = self.torch.empty(max(count, 1), dtype=dtype, device=self.device)
            self.buf[name] = t
        return t

    def _ck(self, rc):
        if rc != 0:
            raise GtbError(self.lib.gtb_esa_error(self.h).decode())

    def begin(self, first_keys):
        torch, dist = self.torch, self.dist
        world, me = dist.get_world_size(), dist.get_rank()
        lo, hi = self.n * me // world, self.n * (me + 1) // world
        cap = hi - lo
        send_k = None if self.positions_only else self._tensor("send_k", cap, torch.int64)
        send_p = self._tensor("send_p", cap, torch.int32)
        fk = np.ascontiguousarray(first_keys, dtype=np.uint64)
        counts = np.zeros(world, dtype=np.uint64)
        self._ck(self.lib.gtb_esa_slice_partition(self.h, self.pl, lo, hi, ptr(fk), world,
                                                  send_k.data_ptr() if send_k is not None else None,
                                                  send_p.data_ptr(), max(cap, 1), ptr(counts)))
        sc = [int(c) for c in counts]
        sct = torch.tensor(sc, dtype=torch.int64, device=self.device)
        rct = torch.empty_like(sct)
        dist.all_to_all_single(rct, sct)
        rc = [int(x) for x in rct.tolist()]
        tot_s, tot_r = sum(sc), sum(rc)
        recv_p = self._tensor("recv_p", tot_r, torch.int32)
        dist.all_to_all_single(recv_p[:tot_r], send_p[:tot_s], output_split_sizes=rc, input_split_sizes=sc)
        if self.positions_only:
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
            self._ck(self.lib.gtb_esa_sort_begin_positions(self.h, self.pl, self.flags, recv_p.data_ptr(), tot_r))
            return tot_s * 4, tot_r * 4
        recv_k = self._tensor("recv_k", tot_r, torch.int64)
        dist.all_to_all_single(recv_k[:tot_r], send_k[:tot_s], output_split_sizes=rc, input_split_sizes=sc)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        self._ck(self.lib.gtb_esa_sort_begin_pairs(self.h, self.pl, self.flags, recv_k.data_ptr(), recv_p.data_ptr(),
                                                   tot_r))
        return tot_s * 12, tot_r * 12


def run_range_distributed(worker, first_keys, dist, device, want_lcp=True, begin=None):
    """lock-step over torch.distributed (NCCL on GPUs, gloo in the CPU tests): rank r of
    the process group runs code range r.  begin (optional) replaces worker.sort_begin()."""
    import torch
    world, me = dist.get_world_size(), dist.get_rank()

    def wait():
        # collectives run on torch's stream, the library on its own: the received data must have
        # landed before the next library kernel reads it
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    if begin is not None:
        begin()
    else:
        worker.sort_begin()
    rounds = 0
    if allmax(worker.unresolved()) > 0:
        worker.ensure_ranks()
        dist.barrier()                                  # every range can answer from here on
        while True:
            # one collective carries the loop condition and the exchange sizes: every rank
            # gathers (positions it asks of each range ..., its own number of tied suffixes)
            send, counts = worker.round_prepare(first_keys, me)        # (nothing to ask when nothing is tied)
            mine = torch.tensor(list(counts) + [worker.unresolved()], dtype=torch.int64, device=device)
            rows = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(rows, mine)
            table = torch.stack(rows).tolist()
            if max(int(row[world]) for row in table) == 0:
                break
            rcl = [int(table[r][me]) for r in range(world)]
            recv_q = torch.empty(sum(rcl), dtype=send.dtype, device=device)
            dist.all_to_all_single(recv_q, send.contiguous(), output_split_sizes=rcl, input_split_sizes=counts)
            wait()
            recv_a = worker.rank_lookup(recv_q)
            back = torch.empty(sum(counts), dtype=send.dtype, device=device)
            dist.all_to_all_single(back, recv_a.contiguous(), output_split_sizes=counts, input_split_sizes=rcl)
            wait()
            worker.round_finish(back)
            rounds += 1
            if rounds > 64:
                raise GtbError("prefix doubling across ranges did not converge")
    worker.sort_end()
    if want_lcp:
        nonempty, first, last = worker.boundary_keys()
        mine = torch.tensor([1 if nonempty else 0, first & 0x7fffffffffffffff, first >> 63,
                             last & 0x7fffffffffffffff, last >> 63], dtype=torch.int64, device=device)
        allk = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        prev = None
        for r in range(me):
            ne, _f, _fh, l, lh = [int(x) for x in allk[r].tolist()]
            if ne:
                prev = l | (lh << 63)
        if nonempty and prev is not None:
            worker.fix_seam(prev)
    return rounds


# ---------------------------------------------------------------- the sharded C entry, one process per GPU
class DistAllgather:
    """The one collective gtb_esa_run_sharded asks of its caller -- an all-gather of a small host
    block -- over torch.distributed (NCCL on the GPUs of a box, gloo in the CPU tests).  This is all
    NCCL carries: the coarse count tables (the count "allreduce": gathered, summed by every rank),
    group sizes, buffer handles and the loop condition of the doubling rounds.  The bulk of a step
    -- partitioned positions, rank lookups -- moves through peer memory inside the kernels."""

    def __init__(self, dist, device):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.world = dist.get_world_size()
        self.cuda = device is not None and device.type == "cuda"
        self.cap = 0
        self.calls = 0
        self.error = None
        self.fn = ALLGATHER_FN(self._call)

    def _reserve(self, nbytes):
        torch = self.torch
        if nbytes <= self.cap:
            return
        cap = max(2 * nbytes, 1 << 16)
        self.h_in = torch.empty(cap, dtype=torch.uint8, pin_memory=self.cuda)
        self.h_out = torch.empty(cap * self.world, dtype=torch.uint8, pin_memory=self.cuda)
        if self.cuda:
            self.d_in = torch.empty(cap, dtype=torch.uint8, device=self.device)
            self.d_out = torch.empty(cap * self.world, dtype=torch.uint8, device=self.device)
        self.cap = cap

    def _call(self, ctx, mine, nbytes, allp):
        try:
            self._reserve(nbytes)
            total = nbytes * self.world
            C.memmove(self.h_in.data_ptr(), mine, nbytes)
            if self.cuda:
                self.d_in[:nbytes].copy_(self.h_in[:nbytes], non_blocking=True)
                self.dist.all_gather_into_tensor(self.d_out[:total], self.d_in[:nbytes])
                self.h_out[:total].copy_(self.d_out[:total], non_blocking=True)
                self.torch.cuda.current_stream(self.device).synchronize()
            else:
                self.dist.all_gather_into_tensor(self.h_out[:total], self.h_in[:nbytes])
            C.memmove(allp, self.h_out.data_ptr(), total)
            self.calls += 1
            return 0
        except Exception as ex:          # never let an exception cross the C frames
            self.error = ex
            return -1


def run_sharded(lib, handle, prefixlength, flags, dist, device, gather=None):
    """rank r of the process group runs code range r of the job: gtb_esa_run_sharded with a NCCL
    all-gather, the buffers of the other ranks mapped through the CUDA virtual-memory API (csrc/gtb_vmm.cuh).  The same C entry the drop-in's
    gtb_group drives with threads."""
    gather = gather or DistAllgather(dist, device)
    rc = lib.gtb_esa_run_sharded(handle, prefixlength, flags, dist.get_rank(), dist.get_world_size(),
                                 gather.fn, None, 1)
    if rc != 0:
        msg = lib.gtb_esa_error(handle).decode()
        if gather.error is not None:
            msg += f" ({gather.error!r})"
        raise GtbError(msg)
    return gather
