#!/usr/bin/env python
"""bench.py -- enhanced-suffix-array construction (`gt suffixerator -suf -lcp -bck`)
on B200: Msuffixes/s for the BASELINE.json workload, the roofline of the dominant
kernel (the onesweep radix pass) and the reference CPU path beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--scale f]
  python bench.py --impl reference ...      # the unmodified reference on host cores

A step = one complete pass of the hot path over the workload: fused key generation +
5-8 radix passes (key length chosen from the text length), group analysis + bucket table
from the sorted keys, text-driven / prefix-doubling refinement of the ties, lcp, special
tail -- results left in HBM (`value`), or through the host-buffer C-ABI including the
H2D copy of the packed sequence and the D2H copy of .suf/.lcp/.llv/.bck (`e2e`: pinned host
buffers, one untimed warm-up step, per-phase breakdown).  At N = 1 the line also carries
`cpu_baseline` (the unmodified reference, one host core, bounded sample of the same generator)
and `cli` (the drop-in binary host/_build/gt_b200 on the same FASTA sample -- FASTA file in, ten index
files out, each compared byte for byte with the reference's; its stage times in `cli.log`; once more with the
reference's FASTA encoder inside the same binary, `cli.seconds_with_reference_encoder`).
With N > 1 ranks (torchrun) the bucket codes are sharded, one range per rank, and every rank calls
the same C entry the drop-in's `gt -j N` runs with threads (gtb_esa_run_sharded): the count gather and
a few small sync blocks travel as NCCL all-gathers, the positions of the position-sharded text scan
(3 ranks and more) and the rank lookups of the doubling rounds go through peer memory inside the
kernels.  GTB_BENCH_PROTOCOL=exchange keeps round 1's NCCL all-to-all request/answer protocol.
`checks.identical_to_reference`: order-dependent checksums of all four tables in HBM, summed over the
ranks, against the checksums of the files the unmodified reference wrote for the workload at this size.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msuffixes/s, suffixerator -suf -lcp"
UNIT = "Msuffixes/s"
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("GTB_BENCH_WORKLOAD", "c4"))
    ap.add_argument("--scale", type=float, default=float(os.environ.get("GTB_BENCH_SCALE", "1.0")))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=64_000_000,
                    help="bases/residues of the CPU baseline sample (about 10-20 s of one host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
                {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------ reference arm / cpu baseline
def reference_sample(workload_name, nsample):
    """a bounded sample of the same workload (same generator, smaller n)"""
    from genometools_b200 import synthetic as sy
    full = {"c2": 100_000_000, "c3": 10_000_000 * 151, "c4": 3_100_000_000, "c5": 500_000_000}[workload_name]
    w = sy.make_workload(workload_name, nsample / full)
    return w


def write_fasta(w, path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import synth
    synth.to_fasta(w.to_symbols(), path, "dna" if w.is_dna else "protein")


def time_reference(w, workdir, runs=1):
    """`gt suffixerator -suf -lcp -bck -pl` of the unmodified reference (oracle/_ref/gtref) --
    single thread: with -j N the reference does not compute lcp values."""
    gtref = os.path.join(ROOT, "oracle", "_ref", "gtref")
    if not os.path.exists(gtref):
        return None
    fa = os.path.join(workdir, "sample.fa")
    write_fasta(w, fa)
    times = []
    for _ in range(runs):
        t0 = time.perf_counter()
        subprocess.check_call([gtref, "suffixerator", "-dna" if w.is_dna else "-protein", "-suf", "-lcp", "-bck",
                               "-pl", "-indexname", os.path.join(workdir, "ref"), "-db", fa],
                              stdout=subprocess.DEVNULL)
        times.append(time.perf_counter() - t0)
    return times


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_reference_sa_only(workdir, is_dna, jobs):
    """`gt -j N suffixerator -suf -bck -pl` -- the reference's multi-threaded
    sorter (src/match/sfx-bentsedg.c:1986-2063) computes the suffix table only (no lcp with -j N)"""
    gtref = os.path.join(ROOT, "oracle", "_ref", "gtref")
    fa = os.path.join(workdir, "sample.fa")
    t0 = time.perf_counter()
    subprocess.check_call([gtref, "-j", str(jobs), "suffixerator", "-dna" if is_dna else "-protein", "-suf", "-bck",
                           "-pl", "-indexname", os.path.join(workdir, "refj"), "-db", fa], stdout=subprocess.DEVNULL)
    return time.perf_counter() - t0


def time_oracle_port(w):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import esa_oracle as eo
    from genometools_b200.suffixerator import recommendedprefixlength
    sym = w.to_symbols()
    t0 = time.perf_counter()
    eo.esa(sym, w.numofchars, recommendedprefixlength(w.numofchars, w.totallength))
    return [time.perf_counter() - t0]


def time_dropin_cli(w, workdir, device):
    """SURVEY.md section 8d, t_cli: the drop-in binary (host/_build/gt_b200 = the reference's `gt` with
    our gt_suffixerator object, FASTA parse and file writes included) on the FASTA sample the reference
    was just timed on, and a byte comparison of the five index files of the two runs."""
    exe = os.path.join(ROOT, "host", "_build", "gt_b200")
    fa = os.path.join(workdir, "sample.fa")
    if not (os.path.exists(exe) and os.path.exists(fa) and os.path.exists(os.path.join(workdir, "ref.suf"))):
        return None
    cmd = [exe, "suffixerator", "-dna" if w.is_dna else "-protein", "-suf", "-lcp", "-bck", "-pl", "-v",
           "-indexname", os.path.join(workdir, "b200"), "-db", fa]

    def once(env=None):
        t0 = time.perf_counter()
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=env).stdout
        dt = time.perf_counter() - t0
        notes = [ln[2:] for ln in out.split("\n") if ln.startswith("# B200 encoder") or ln.startswith("# wall seconds")]
        return dt, notes

    runs, notes = [], []
    for _ in range(3):          # the first start of the binary on a fresh box pages in the CUDA driver
        dt, notes = once()
        runs.append(dt)
    t = min(runs)
    same = True
    for ext in ("suf", "lcp", "llv", "bck", "prj", "esq", "ssp", "des", "sds", "md5"):
        a, b = os.path.join(workdir, "ref." + ext), os.path.join(workdir, "b200." + ext)
        if os.path.exists(a) or os.path.exists(b):
            same = same and subprocess.call(["cmp", "-s", a, b]) == 0
    # the same binary with the reference's one-core FASTA encoder in front of the sorter
    t_refenc, _ = once(dict(os.environ, GTB200_ENCODER="reference"))
    return {"seconds": t, "seconds_runs": runs, "value": (w.totallength + 1) / t / 1e6, "unit": UNIT,
            "files_identical_to_reference": same, "log": notes,
            "seconds_with_reference_encoder": t_refenc,
            "what": "host/_build/gt_b200 suffixerator -suf -lcp -bck -pl on the cpu_baseline FASTA sample: the "
                    "reference's own CLI, option parser, loader and .prj writer around libgtb200 (process start, "
                    "CUDA context, FASTA -> .esq/.ssp/.des/.sds/.md5 by gtb_fasta_encode on the host cores and the "
                    "writes of .suf/.lcp/.llv/.bck included); all ten index files compared with the reference's"}


def cpu_baseline(args, wl_name, device=None):
    w = reference_sample(wl_name, args.cpu_sample)
    tmp = tempfile.mkdtemp(prefix="gtb_ref_")
    cli, sa_only = None, None
    try:
        times = time_reference(w, tmp)
        kind = "reference"
        if times is None:
            times = time_oracle_port(w)
            kind = "port"
        else:
            jobs = os.cpu_count() or 1
            try:
                tj = time_reference_sa_only(tmp, w.is_dna, jobs)
                sa_only = {"value": (w.totallength + 1) / tj / 1e6, "unit": "Msuffixes/s (suffix table only)",
                           "threads": jobs, "seconds": tj,
                           "command": f"gtref -j {jobs} suffixerator -suf -bck -pl (no -lcp: the reference's threaded "
                                      "sorter writes no lcp values)"}
            except Exception as ex:
                sa_only = {"error": str(ex)}
            if device is not None:
                try:
                    cli = time_dropin_cli(w, tmp, device)
                except Exception as ex:      # reported, never fatal
                    cli = {"error": str(ex)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    t = min(times)
    return {"value": (w.totallength + 1) / t / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
            "seconds": t, "cpu_model": cpu_model(), "host_cores": os.cpu_count(),
            "sa_only_all_cores": sa_only,
            "sample": f"{w.name} generator at n={w.totallength} ({w.description}); "
                      + ("oracle/_ref/gtref suffixerator -suf -lcp -bck -pl, 1 thread (lcp is wrong with -j N)"
                         if kind == "reference" else "oracle/esa_oracle.c restatement")}, cli, w


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl = args.workload
    w = reference_sample(wl, args.cpu_sample)
    tmp = tempfile.mkdtemp(prefix="gtb_ref_")
    try:
        fa = os.path.join(tmp, "sample.fa")
        gtref = os.path.join(ROOT, "oracle", "_ref", "gtref")
        kind = "reference" if os.path.exists(gtref) else "port"
        times = []
        for it in range(args.warmup + args.steps):
            if kind == "reference":
                t = time_reference(w, tmp)[0]
            else:
                t = time_oracle_port(w)[0]
            if it >= args.warmup:
                times.append(t)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    tot = sum(times)
    val = (w.totallength + 1) * len(times) / tot / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": wl, "sample_totallength": w.totallength, "description": w.description,
                       "command": "suffixerator -suf -lcp -bck -pl (1 thread: the reference computes no lcp "
                                  "values with -j N)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "cpu_model": cpu_model(),
                             "host_cores": os.cpu_count(),
                             "sample": f"{w.name} generator at n={w.totallength}, every step (a bounded sample of the "
                                       "workload; the B200 arm reports `same_input` on exactly this sample)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def same_input_run(lib, device, w, cpu, steps=5):
    """value and e2e of the B200 path on the bounded sample the reference arm sorts: a ratio on ONE input"""
    import torch
    from genometools_b200._lib import GtbStats, ptr
    from genometools_b200.suffixerator import recommendedprefixlength
    n = w.totallength
    pl = recommendedprefixlength(w.numofchars, n)
    buf = C.create_string_buffer(512)
    h = lib.gtb_esa_new(device, buf, 512)
    if not h:
        return {"error": buf.value.decode()}
    try:
        def ck(rc):
            if rc != 0:
                raise RuntimeError(lib.gtb_esa_error(h).decode())

        def upload():
            if w.is_dna:
                ck(lib.gtb_esa_set_input_2bit(h, ptr(w.words), w.words.shape[0], n,
                                              ptr(w.ranges) if w.ranges.shape[0] else None, w.ranges.shape[0]))
            else:
                ck(lib.gtb_esa_set_input_bytes(h, ptr(w.symbols), n, w.numofchars))
        upload()
        stream = torch.cuda.ExternalStream(lib.gtb_esa_stream(h), device=torch.device("cuda", device))
        for _ in range(2):
            ck(lib.gtb_esa_run(h, pl, 7))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            ck(lib.gtb_esa_run(h, pl, 7))
        e1.record(stream)
        torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) / 1e3 / steps
        e = n + 1
        suf = torch.empty(e, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        lcp = torch.empty(e, dtype=torch.uint8, pin_memory=True).numpy()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(w.numofchars, pl, C.byref(a), C.byref(b), C.byref(c))
        lb = np.empty(a.value + 1, dtype=np.uint32); csc = np.empty(max(b.value, 1), dtype=np.uint32)
        dist = np.empty(max(c.value, 1), dtype=np.uint32)
        times = []
        for it in range(3):
            t0 = time.perf_counter()
            upload()
            ck(lib.gtb_esa_run(h, pl, 7))
            k = int(lib.gtb_esa_num_llv(h))
            llv = np.empty(2 * max(k, 1), dtype=np.uint64)
            ck(lib.gtb_esa_copy_results(h, ptr(suf), ptr(lcp), ptr(llv) if k else None, ptr(lb), ptr(csc),
                                        ptr(dist) if c.value else None))
            times.append(time.perf_counter() - t0)
        e2e_s = min(times[1:])
        ref = cpu["value"] if cpu and cpu.get("value") else None
        out = {"input": f"{w.name} generator at n={n}: the cpu_baseline sample, identical on both sides",
               "b200_value": (n + 1) / dev_s / 1e6, "b200_e2e": (n + 1) / e2e_s / 1e6, "reference_value": ref,
               "unit": UNIT, "b200_ms_per_step": dev_s * 1e3, "b200_e2e_ms": e2e_s * 1e3}
        if ref:
            out["ratio_value"] = out["b200_value"] / ref
            out["ratio_e2e"] = out["b200_e2e"] / ref
        return out
    finally:
        lib.gtb_esa_delete(h)


# ------------------------------------------------------------------ the B200 arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    from genometools_b200 import _lib, synthetic as sy
    from genometools_b200._lib import GtbStats, GTB_WANT_SUF, GTB_WANT_LCP, GTB_WANT_BCK, GTB_REUSE_COUNTS, ptr
    from genometools_b200.suffixerator import recommendedprefixlength
    from genometools_b200.sharding import suftab_parts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = _lib.load()
    w = sy.make_workload(args.workload, args.scale)
    n = w.totallength
    pl = recommendedprefixlength(w.numofchars, n)
    buf = C.create_string_buffer(512)
    h = lib.gtb_esa_new(local_rank, buf, 512)
    if not h:
        raise SystemExit("gtb_esa_new: " + buf.value.decode())

    def ck(rc):
        if rc != 0:
            raise SystemExit("libgtb200: " + lib.gtb_esa_error(h).decode())

    def upload():
        if w.is_dna:
            ck(lib.gtb_esa_set_input_2bit(h, ptr(w.words), w.words.shape[0], n,
                                          ptr(w.ranges) if w.ranges.shape[0] else None, w.ranges.shape[0]))
        else:
            ck(lib.gtb_esa_set_input_bytes(h, ptr(w.symbols), n, w.numofchars))

    upload()
    flags = GTB_WANT_SUF | GTB_WANT_LCP | GTB_WANT_BCK
    # weak scaling: the sequence is replicated, the bucket codes are sharded over the ranks.
    # Every step: count allreduce (each rank counts coarse codes over 1/world of the text), the
    # same code ranges cut on every rank, then the lock-step sort of the rank's own range.
    dev = torch.device("cuda", local_rank)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    st = GtbStats()
    protocol = os.environ.get("GTB_BENCH_PROTOCOL", "sharded")     # "exchange": the round-1 NCCL request/answer protocol
    worker = gather = None
    if world > 1:
        from genometools_b200.multirange import (GpuRangeWorker, run_range_distributed, range_first_keys,
                                                 coarse_allreduce_and_split, allreduce_bcktab, PairExchange,
                                                 DistAllgather, run_sharded)
        # the host threads that widen the suffix table are shared by the ranks of the box
        os.environ.setdefault("GTB200_HOST_THREADS", str(max(1, ((os.cpu_count() or 4) - 2) // world)))
        if protocol == "sharded":
            gather = DistAllgather(dist, dev)
        else:
            worker = GpuRangeWorker(h, pl, flags | GTB_REUSE_COUNTS, local_rank)
            # sharding the text scan pays once the all-to-all is small against the scans it saves
            use_x = os.environ.get("GTB_BENCH_EXCHANGE", "1" if world >= 4 else "0") == "1"
            exchange = PairExchange(lib, h, pl, flags | GTB_REUSE_COUNTS, n, dist, dev) if use_x else None

    def step():
        if world > 1 and protocol == "sharded":
            # the C entry the drop-in's `gt -j N` runs with threads, here one process per GPU: NCCL carries
            # the count gather and the small sync blocks, peer memory (mapped through the CUDA virtual-memory API)
            # everything heavy
            run_sharded(lib, h, pl, flags, dist, dev, gather)
        elif world > 1:
            parts = coarse_allreduce_and_split(lib, h, pl, n, dist, dev)
            if len(parts) != world:
                raise SystemExit("could not cut the bucket codes into one part per rank")
            mn, mx, off, width = parts[rank]
            ck(lib.gtb_esa_set_code_range_known(h, mn, mx, off, width, 1 if rank == world - 1 else 0))
            fk = range_first_keys(w.numofchars, pl, parts)
            run_range_distributed(worker, fk, dist, dev,
                                  begin=(lambda: exchange.begin(fk)) if exchange is not None else None)
        else:
            ck(lib.gtb_esa_run(h, pl, flags))
        ck(lib.gtb_esa_get_stats(h, C.byref(st)))
        bck_state["summed"] = False
        return st.as_dict()

    bck_state = {"summed": False}

    def sum_bck():
        # a rank's run fills the bucket-table entries of its own codes: the job's table is their sum
        if dist is not None and not bck_state["summed"]:
            allreduce_bcktab(lib, h, w.numofchars, pl, dist, dev)
            bck_state["summed"] = True

    lib_stream = torch.cuda.ExternalStream(lib.gtb_esa_stream(h), device=dev)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)                      # CUDA events on the stream the kernels are launched on
    t0 = time.perf_counter()
    radix_ms, hist_ms, launches, radix_passes, pairs = 0.0, 0.0, 0, 0, 0
    first_ms, first_passes, first_pairs = 0.0, 0, 0
    last = None
    for _ in range(args.steps):
        last = step()
        radix_ms += last["ms_radix"]; hist_ms += last["ms_hist"]
        launches += last["kernel_launches"]; radix_passes += last["radix_passes"]; pairs += last["radix_pairs_moved"]
        first_ms += last["ms_radix_first"]; first_passes += last["radix_passes_first"]; first_pairs += last["radix_pairs_first"]
    ev1.record(lib_stream)
    sync_all()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.result()
    # device time of the K steps, max over ranks
    tm = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = tm.tolist()
    total_units = (n + 1) * args.steps          # all ranks together sort every suffix once per step
    value = total_units / (dev_ms_max / 1e3) / 1e6

    # ---- roofline of the dominant kernel (rs_onesweep_kernel), measured live (ms_radix =
    # CUDA events around the pass launches on the library's stream) ----
    N = last["nonspecials"]
    # The graded launches are the passes of the FIRST-LEVEL sort (all suffixes of the rank per launch: the
    # dominant kernel at its dominant size).  Every such pass moves 24 B per pair (12 read + 12 written)
    # except the first of a step, which reads the 2-bit text (n/4) and the special mask (n/8) instead of
    # 12 B per pair (a rank of a sharded scan reads its slice of the text in the partition pass).
    text_bytes = 0.375 * n / world
    alg_bytes = 24.0 * first_pairs - args.steps * 12.0 * N + args.steps * text_bytes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (first_ms / 1e3) / 1e9 if first_ms > 0 else 0.0
    # all launches of the kernel, the small passes of the refinement rounds included (launch-latency bound)
    alg_all = 24.0 * pairs - args.steps * 12.0 * N + args.steps * text_bytes
    achieved_all = alg_all / (radix_ms / 1e3) / 1e9 if radix_ms > 0 else 0.0
    traffic = None          # DRAM bytes per launch: ncu's bytes per pair (profiles/) x pairs per launch here
    tpath = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.exists(tpath) and first_passes:
        try:
            traffic = float(json.load(open(tpath))["dram_bytes_per_pair"]) * first_pairs / first_passes
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "rs_onesweep_kernel (one 8-bit LSD pass over (key64,pos32) pairs), the passes "
                                          "of the first-level sort",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": (alg_bytes / first_passes) if first_passes else None,
                "launches_timed": first_passes, "ms_per_launch": first_ms / first_passes if first_passes else None,
                "share_of_step": first_ms / dev_ms if dev_ms else None,
                "all_launches": {"launches": radix_passes, "achieved": achieved_all,
                                 "frac": achieved_all / peak if peak else None,
                                 "ms_per_launch": radix_ms / radix_passes if radix_passes else None,
                                 "share_of_step": radix_ms / dev_ms if dev_ms else None,
                                 "note": "incl. the refinement rounds' passes over a few 10^4..10^7 ties each"}}

    # ---- end to end through the host-buffer API: H2D of the packed sequence, all kernels,
    # D2H of suftab (uint64), lcptab, llv, bucket table ----
    e2e = None
    if args.e2e_steps > 0:
        ent_max = int(lib.gtb_esa_num_entries(h)) + 16      # this rank's share of suftab/lcptab
        suf = torch.empty(ent_max, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        lcp = torch.empty(ent_max, dtype=torch.uint8, pin_memory=True).numpy()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(w.numofchars, pl, C.byref(a), C.byref(b), C.byref(c))
        lbh = torch.empty(a.value + 1, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        csch = torch.empty(max(b.value, 1), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        disth = torch.empty(max(c.value, 1), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        if w.is_dna:
            pw = torch.empty(w.words.shape[0], dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
            pw[:] = w.words
            words_pinned = pw
        else:
            ps = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
            ps[:] = w.symbols
        llv_pinned = torch.empty((max(int(lib.gtb_esa_num_llv(h)) * 2, 1024), 2), dtype=torch.int64,
                                 pin_memory=True).numpy().view(np.uint64)
        d2h = 0
        # gtb_esa_run_to_host (the suffix table leaves for the host while the ties are refined) is correct
        # (tests/test_gpu_configs.py) but does not pay on the bench hosts: 532 ms against 522 ms for c4 --
        # measured, profiles/README.md -- so the timed path is run + gtb_esa_copy_results
        overlap_copy = os.environ.get("GTB_BENCH_OVERLAP", "0") == "1"
        parts_s = [0.0, 0.0, 0.0]                   # H2D, kernels, D2H of the timed steps
        d2h_parts = [0.0]                           # gtb_esa_copy_results (incl. the warm-up step)

        def e2e_step():
            nonlocal d2h
            ta = time.perf_counter()
            if w.is_dna:
                ck(lib.gtb_esa_set_input_2bit(h, ptr(words_pinned), words_pinned.shape[0], n,
                                              ptr(w.ranges) if w.ranges.shape[0] else None, w.ranges.shape[0]))
            else:
                ck(lib.gtb_esa_set_input_bytes(h, ptr(ps), n, w.numofchars))
            tb = time.perf_counter()
            if world == 1 and overlap_copy:
                # one call: the suffix table crosses PCIe while the ties are refined (gtb_esa_run_to_host)
                nl = C.c_uint64()
                ck(lib.gtb_esa_run_to_host(h, pl, flags, ptr(suf), ptr(lcp), ptr(llv_pinned), llv_pinned.shape[0],
                                           C.byref(nl), ptr(lbh), ptr(csch), ptr(disth) if c.value else None))
                ck(lib.gtb_esa_get_stats(h, C.byref(st)))
                bck_state["summed"] = False
                e, k = lib.gtb_esa_num_entries(h), int(nl.value)
                d2h = 8 * e + e + 16 * k + 4 * (a.value + 1 + b.value + c.value)
                td = time.perf_counter()
                d2h_parts[0] += td - tb
                return tb - ta, 0.0, td - tb
            step()
            tc = time.perf_counter()
            e = lib.gtb_esa_num_entries(h)
            t1 = time.perf_counter()
            k = lib.gtb_esa_num_llv(h)
            if k > llv_pinned.shape[0]:
                raise SystemExit("bench.py: llv buffer too small")
            # the gather: every rank copies its shard (its part of .suf/.lcp/.llv); the bucket table of the
            # job is the sum of the ranks' tables (NCCL all-reduce) and leaves from rank 0.
            # One call: lcptab, llv and the bucket table travel beside the suffix table
            with_bck = rank == 0
            sum_bck()
            ck(lib.gtb_esa_copy_results(h, ptr(suf), ptr(lcp), ptr(llv_pinned) if k else None,
                                        ptr(lbh) if with_bck else None, ptr(csch) if with_bck else None,
                                        ptr(disth) if (c.value and with_bck) else None))
            d2h_parts[0] += time.perf_counter() - t1
            # bytes of the host tensors that are filled (uint64 suftab, uint8 lcptab, llv, bucket
            # tables).  Fewer cross the bus: part of the suffix table travels as uint32 and is
            # widened by host threads, the rest is widened on the device (gtb_esa_copy_suftab_u64)
            d2h = 8 * e + e + 16 * k + (4 * (a.value + 1 + b.value + c.value) if with_bck else 0)
            td = time.perf_counter()
            return tb - ta, tc - tb, td - tc

        e2e_step()                                  # untimed: pins the staging buffers of the copies
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            for i_, v_ in enumerate(e2e_step()):
                parts_s[i_] += v_
        sync_all()
        ewall = time.perf_counter() - t0
        te = torch.tensor([ewall], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        tb_ = torch.tensor([int(d2h)], dtype=torch.int64, device="cuda")
        if dist is not None:
            dist.all_reduce(tb_)
        e2e = {"value": (n + 1) * args.e2e_steps / te.item() / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": w.input_bytes() * world, "d2h_bytes_per_step": int(tb_.item()),
               "steps": args.e2e_steps, "ms_per_step": te.item() / args.e2e_steps * 1e3,
               "breakdown_ms": ({"h2d": parts_s[0] / args.e2e_steps * 1e3,
                                 "run_to_host (kernels and copies overlapped)": parts_s[2] / args.e2e_steps * 1e3}
                                if (world == 1 and overlap_copy) else
                                {"h2d": parts_s[0] / args.e2e_steps * 1e3, "kernels": parts_s[1] / args.e2e_steps * 1e3,
                                 "d2h": parts_s[2] / args.e2e_steps * 1e3,
                                 "d2h_all_tables_one_call": d2h_parts[0] / (args.e2e_steps + 1) * 1e3}),
               "api": "gtb_esa_set_input_* + gtb_esa_run_to_host" if (world == 1 and overlap_copy) else
                      "gtb_esa_set_input_* + run + gtb_esa_copy_results",
               "note": "pinned host buffers; results delivered as the files hold them (uint64 suftab, "
                       "uint8 lcptab, llv pairs, uint32 bucket tables); one untimed warm-up step"}

    # cheap global check: the shards together are a permutation of 0..n (sum of entries)
    class _DevArr:
        def __init__(self, p, count, typestr):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (p, False), "version": 2}
    ent = lib.gtb_esa_num_entries(h)
    sa_t = torch.as_tensor(_DevArr(lib.gtb_esa_dev_suftab(h), ent, "<u4"), device=dev) if ent else None
    ssum = int(sa_t.to(torch.int64).sum().item()) if ent else 0
    chk = torch.tensor([ssum, ent, last["nonspecials"], int(last["lcptabsum"]), last["numoflargelcpvalues"]],
                       dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(chk)
    mx = torch.tensor([last["maxbranchdepth"], last["doubling_rounds"]], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    bck_ok = None
    if dist is not None:
        # the ranks' bucket tables hold their own codes only: sum them and check the total
        sum_bck()
        a_, b_, c_ = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.gtb_bck_sizes(w.numofchars, pl, C.byref(a_), C.byref(b_), C.byref(c_))
        plb_ = C.c_void_p()
        lib.gtb_esa_dev_bcktab(h, C.byref(plb_), None, None)
        lbt = torch.as_tensor(_DevArr(plb_.value, a_.value + 1, "<u4"), device=dev)
        bck_ok = int(lbt[-1].item()) == int(chk[2].item()) and bool((lbt[1:].to(torch.int64) >= lbt[:-1].to(torch.int64)).all().item())
    # order-dependent checksums of the tables in HBM (gtb_esa_hash_results; shards add up) against the
    # checksums of the files the unmodified reference wrote for this workload at this size
    # (tests/golden/config_md5.json, tests/golden/make_golden_configs.py)
    def as_i64(v):
        return v - (1 << 64) if v >= (1 << 63) else v
    nllv_mine = int(lib.gtb_esa_num_llv(h))
    llv_before = 0
    if dist is not None:
        allk = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allk, torch.tensor([nllv_mine], dtype=torch.int64, device=dev))
        llv_before = sum(int(t.item()) for t in allk[:rank])
    out3 = (C.c_uint64 * 3)()
    ck(lib.gtb_esa_hash_results(h, llv_before, out3))
    hs = torch.tensor([as_i64(int(out3[i])) for i in range(3)], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(hs)             # (int64 sums wrap: the checksum is a sum mod 2^64)
    hb = C.c_uint64()
    ck(lib.gtb_esa_hash_bcktab(h, C.byref(hb)))       # (N > 1: the tables were summed over the ranks above)
    hashes = {"suf": int(hs[0].item()) % (1 << 64), "lcp": int(hs[1].item()) % (1 << 64),
              "llv": int(hs[2].item()) % (1 << 64), "bck": int(hb.value)}
    golden_key = args.workload if args.scale == 1.0 else f"{args.workload}@{args.scale:g}"
    gpath = os.path.join(ROOT, "tests", "golden", "config_md5.json")
    golden = json.load(open(gpath)).get(golden_key) if os.path.exists(gpath) else None
    identical = None
    if golden is not None and golden["totallength"] == n:
        identical = all(hashes[e] == golden["files"][e]["mixhash"] for e in ("suf", "lcp", "llv", "bck"))
    checks = {"identical_to_reference": identical,
              "identical_to_reference_how": "order-dependent 64-bit checksums (genometools_b200/mixhash.py) of "
              "suftab/lcptab/llv/bucket table in HBM, summed over the shards, equal to the checksums of the "
              "files `gtref suffixerator` wrote for this workload (tests/golden/config_md5.json)"
              if golden is not None else "no golden checksums for this workload/scale",
              "suf_hash": hashes["suf"], "lcp_hash": hashes["lcp"], "llv_hash": hashes["llv"], "bck_hash": hashes["bck"],
              "suftab_is_permutation_sum": int(chk[0].item()) == n * (n + 1) // 2 and int(chk[1].item()) == n + 1,
              "lcptabsum": int(chk[3].item()), "largelcpvalues": int(chk[4].item()),
              "maxbranchdepth": int(mx[0].item()), "doubling_rounds": int(mx[1].item())}
    if bck_ok is not None:
        checks["merged_bucket_table_monotone_and_complete"] = bck_ok
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu, cli, same_input = None, None, None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu, cli, wsample = cpu_baseline(args, args.workload, local_rank)
            # the same input on both sides: the B200 path on exactly the sample the reference just sorted
            # (device time with the sequence resident, and end to end through host buffers)
            same_input = same_input_run(lib, local_rank, wsample, cpu)
        except Exception as ex:      # the baseline is reported, never fatal
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": str(ex)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "description": w.description, "totallength": n,
                   "specialcharacters": int(last["specialcharacters"]), "prefixlength": pl,
                   "numofchars": w.numofchars, "outputs": "-suf -lcp -bck (results resident in HBM; `value` excludes the H2D copy of the packed "
                              "sequence -- inputs resident when the timed region starts -- `e2e` includes it)",
                   "dtype_note": "u32 positions in HBM (n + 1 < 2^32), widened to the file's uint64 on copy-out",
                   "l2": "inputs_exceed_l2 (no flush needed)", "sharding": f"{world} bucket-code ranges, one per GPU "
                   f"({protocol}: " + ("gtb_esa_run_sharded -- NCCL all-gathers of small host blocks, positions and "
                                       "rank lookups through peer memory (cuMemCreate allocations mapped by the peers)" if protocol == "sharded" else
                                       "NCCL all-to-all request/answer protocol") + ")"
                   if world > 1 else "single range", "scale": args.scale},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "same_input": same_input, "cli": cli,
        "wall_ms_per_step": wall_ms_max / args.steps,
        "breakdown_ms_last_step": {k: last[k] for k in ("ms_count", "ms_hist", "ms_radix", "ms_analyze",
                                                         "ms_doubling", "ms_lcp", "ms_tail")},
        "checks": checks,
        "last_step": {k: int(last[k]) for k in ("unresolved_after_first_sort", "doubling_rounds", "radix_passes",
                                                 "maxbranchdepth", "numoflargelcpvalues", "longest")},
    }
    print(json.dumps(line))
    lib.gtb_esa_delete(h)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
