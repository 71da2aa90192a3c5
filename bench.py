#!/usr/bin/env python
"""bench.py -- events/s through divide_rounds + decide_fame (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3]

A "step" is one pass of the hot path over the whole synthetic trace of the
workload: for every chunk of the pinned call schedule, `divide_rounds(chunk)` then
`decide_fame()` (swirld.py:325-327), starting from an engine whose consensus state
was cleared.  Workloads (BASELINE.json `configs`):
    c1  4 members,  2 000 events, K=50      (plumbing)
    c2  16 members, 100 000 events, K=4096
    c3  64 members, 1 000 000 events, K=65536   <- headline, default
Trace: generator G1 (reference-sim gossip), seed 1 + rank.

`value`   : events/s with the event columns already resident in HBM (sw_rewind keeps
            them), device-timed with CUDA events on the engine's stream.
`e2e`     : same metric through the C ABI from HOST buffers: every step clears the
            engine, appends each chunk from pinned host memory (H2D inside the timed
            region), runs divide_rounds + decide_fame per chunk and reads back
            round / witness / famous for all events (D2H inside the timed region).
`roofline`: the dominant kernel k_rounds_batch (round numbers of a chunk on the whole GPU, one
            cooperative launch per divide_rounds call): algorithmic bytes 4M + 8 + 5 + M/8 per
            event (read the event's can_see row, p0 and creator; write round, witness flag and
            seen-mask: SURVEY.md section 8d's B(M) minus the can_see part B1(M), plus the row
            read that a separate kernel cannot avoid) over its mean launch duration (CUDA
            events on the engine's stream), against MEASURED_PEAKS.json hbm_gbs.
`roofline_can_see`: the same for the can_see kernel family k_cs_* (SURVEY.md section 8d's
            B1(M) = 12M + 12 bytes per event), the bandwidth-bound part of the path.
`cpu_baseline` / --impl reference: the reference is pure Python and cannot travel to
            the GPU box, so the CPU arm is the literal C restatement oracle/
            (kind "port"), single-threaded per node-view like the reference
            (README.md:27-28); at N > 1 it runs the N replicas of the GPU arm's workload in
            N parallel processes (cores = N).

N > 1: one process per GPU (torchrun), each rank runs an independent node-view
(its own trace, seed 1 + rank) -- the path has no cross-GPU exchange at M <= 64
("replicas"), so scaling is weak and there is no collective
on the data path; ranks meet at a barrier before and after the timed region and
the time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "py-swirld_b200"))

import numpy as np  # noqa: E402

WORKLOADS = {
    "c1": dict(M=4, N=2000, K=50),
    "c2": dict(M=16, N=100000, K=4096),
    "c3": dict(M=64, N=1000000, K=65536),
}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at c3 (65536 events), from the committed
# ncu --set full captures under profiles/ (see profiles/README.md)
ROUNDS_DRAM_BYTES_PER_LAUNCH = 18547712 + 2816        # k_rounds_batch (profiles/r01c_ncu_full.md)
WALKER_DRAM_BYTES_PER_LAUNCH = 2365952 + 2083072      # k_divide_levels (SW_DIVIDE_IMPL=4)


def algorithmic_bytes_per_event(M):
    """divide_rounds + decide_fame combined."""
    return 12 * M + 12 + 5 + M / 8.0


def can_see_bytes_per_event(M):
    """read two parent rows, write one, read p0/p1/creator."""
    return 12 * M + 12


def rounds_bytes_per_event(M):
    """k_rounds_batch: read row(h), p0, creator; write round, witness flag, seen-mask."""
    return 4 * M + 8 + 5 + M / 8.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) >= 6 and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_trace(wl, seed):
    from swirld_b200 import traces
    cache = "/tmp/swirld_trace_M%d_N%d_s%d.npz" % (wl["M"], wl["N"], seed)
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return traces.Trace(wl["M"], z["p0"], z["p1"], z["creator"], z["t"], z["sig"], "G1 cached")
        except Exception:
            pass
    tr = traces.gossip(wl["M"], wl["N"], seed)
    try:
        np.savez(cache, p0=tr.p0, p1=tr.p1, creator=tr.creator, t=tr.t, sig=tr.sig)
    except Exception:
        pass
    return tr


# --------------------------------------------------------------------------- CPU arm
def run_cpu_pass(tr, K, limit=None):
    """One pass of divide_rounds + decide_fame (+ find_order, untimed) through the
    oracle port; returns (events, seconds in dr+df)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    n = tr.N if limit is None else min(limit, tr.N)
    res = orc.run_oracle(tr.slice(0, n), K)
    res["oracle"].close()
    return n, res["t_divide_rounds"] + res["t_decide_fame"], res["t_find_order"]


def reference_sample_events(wl, steps):
    """Events per step of the CPU arm: the whole trace when the run stays within about a minute of CPU work
    (the port does ~2e5 events/s on one core), else a prefix of whole chunks."""
    budget = int(60 * 2.0e5 / max(1, steps))
    if budget >= wl["N"]:
        return wl["N"]
    return max(wl["K"], budget // wl["K"] * wl["K"])


def _reference_replica(job):
    """One node-view through the CPU port (runs in its own process when world > 1)."""
    wl, seed, steps, warm = job
    tr = make_trace(wl, seed)
    limit = reference_sample_events(wl, steps)
    if warm:
        run_cpu_pass(tr, wl["K"], limit=min(tr.N, 50000))
    t0 = time.perf_counter()
    tot_e, tot_s, t_fo = 0, 0.0, 0.0
    for _ in range(steps):
        n, s, fo = run_cpu_pass(tr, wl["K"], limit)
        tot_e += n
        tot_s += s
        t_fo += fo
    return tot_e, tot_s, t_fo, time.perf_counter() - t0


def bench_reference(args, wl, rank, world):
    """The CPU arm on the workload of the GPU arm: `world` independent node-views (replicas), one host core
    each -- the reference is single-threaded per node (README.md:27-28), so replicas are the only way it
    can use more cores."""
    if rank != 0:
        return
    jobs = [(wl, rank_seed(r), args.steps, args.warmup >= 1) for r in range(world)]
    if world == 1:
        res = [_reference_replica(jobs[0])]
    else:
        import concurrent.futures as cf
        import multiprocessing as mp
        with cf.ProcessPoolExecutor(world, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_reference_replica, jobs))
    tot_e = sum(r[0] for r in res)
    slow = max(r[1] for r in res)                # the job is as slow as its slowest replica
    t_fo = max(r[2] for r in res)
    v = tot_e / slow
    line = {
        "impl": "reference", "metric": "events/sec divide_rounds+decide_fame", "value": v, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * slow / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(wl, world),
        "cpu_baseline": {"value": v, "unit": "events/s", "cores": world, "kind": "port",
                         "sample": "first %d of %d events x %d passes per replica, %d replica(s) in parallel processes, "
                                   "oracle/swirld_oracle.c (literal C restatement; the Python reference cannot travel to "
                                   "the GPU box), host has %d cpus" % (reference_sample_events(wl, args.steps), wl["N"],
                                                                      args.steps, world, os.cpu_count())},
        "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "find_order_events_per_s": tot_e / t_fo if t_fo > 0 else None,
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, world):
    return {"workload": "G1 reference-sim gossip trace, %d members x %d events, call schedule K=%d events per "
                        "(divide_rounds, decide_fame) pair, unit stake, coin period 6" % (wl["M"], wl["N"], wl["K"]),
            "members": wl["M"], "events": wl["N"], "chunk": wl["K"],
            "parallelism": "replicas x%d (independent node-views, no collective)" % world,
            "l2": "L2 flushed (256 MiB device memset) before every timed step; per-step working set "
                  "(can_see + T tables) also exceeds L2 at c3"}


# --------------------------------------------------------------------------- multi-rank plumbing
def rank_seed(rank):
    """Every rank advances its own node-view: an independent trace."""
    return 1 + rank


def max_over_ranks(values, dist, device):
    """Element-wise max of a few per-rank timings (the job is as slow as its slowest rank)."""
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def whole_job_rate(world, events_per_rank_step, steps, ms_max):
    """events/s of the whole job: all ranks' events over the slowest rank's time."""
    return world * events_per_rank_step * steps / (ms_max * 1e-3)


# --------------------------------------------------------------------------- GPU arm
def bench_ours(args, wl, rank, world, local_rank):
    import torch
    from swirld_b200 import engine
    from swirld_b200.traces import chunks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    M, N, K = wl["M"], wl["N"], wl["K"]
    tr = make_trace(wl, rank_seed(rank))
    # pinned host copies of the event columns (the e2e leg's source)
    pin = {}
    for k in ("p0", "p1", "creator", "t", "sig"):
        src = torch.from_numpy(np.ascontiguousarray(getattr(tr, k)))
        pin[k] = src.pin_memory().numpy()
    out_round = torch.empty(N, dtype=torch.int32).pin_memory().numpy()
    out_wit = torch.empty(N, dtype=torch.uint8).pin_memory().numpy()
    out_fam = torch.empty(N, dtype=torch.int8).pin_memory().numpy()

    eng = engine.Engine(M, N, device=local_rank)
    sched = list(chunks(N, K))

    def append_all():
        for first, cnt in sched:
            s = slice(first, first + cnt)
            eng.append(pin["p0"][s], pin["p1"][s], pin["creator"][s], pin["t"][s], pin["sig"][s])

    timed_launches = [0]

    def step_resident():
        eng.rewind()
        eng.flush_l2()
        l0 = eng.stats()["kernel_launches"]       # (syncs; before the timed region)
        eng.record(0)
        for first, cnt in sched:
            eng.divide_rounds(first, cnt)
            eng.decide_fame()
        eng.record(1)
        ms = eng.elapsed_ms(0, 1)
        timed_launches[0] += eng.stats()["kernel_launches"] - l0
        return ms

    def step_e2e():
        eng.reset()
        eng.flush_l2()
        eng.sync()
        t0 = time.perf_counter()

        def feed(i):
            first, cnt = sched[i]
            s = slice(first, first + cnt)
            eng.append(pin["p0"][s], pin["p1"][s], pin["creator"][s], pin["t"][s], pin["sig"][s])

        # the caller's software pipeline: appends run two chunks ahead of the consensus calls (sw_append: host
        # checks, asynchronous H2D and the can_see scan of the new events on the engine's copy stream), so
        # they overlap the round kernels of earlier chunks; sw_decide_fame is the synchronising call
        ahead = 2
        for i in range(min(ahead, len(sched))):
            feed(i)
        for i, (first, cnt) in enumerate(sched):
            eng.divide_rounds(first, cnt)
            if i + ahead < len(sched):
                feed(i + ahead)
            eng.decide_fame()
        lib, h = eng._lib, eng._h
        import ctypes as C
        lib.sw_get_round(h, 0, N, out_round.ctypes.data_as(C.c_void_p))
        lib.sw_get_witness_flags(h, 0, N, out_wit.ctypes.data_as(C.c_void_p))
        lib.sw_get_famous(h, 0, N, out_fam.ctypes.data_as(C.c_void_p))
        eng.sync()
        return (time.perf_counter() - t0) * 1e3

    # ---- resident-input leg (the contract's timed region)
    append_all()
    for _ in range(args.warmup):
        step_resident()
    st0 = eng.stats()
    timed_launches[0] = 0
    barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(args.steps):
            dev_ms += step_resident()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
    st1 = eng.stats()
    check = (int(eng.rounds().astype(np.int64).sum()), int(eng.max_round), int(len(eng.consensus())))

    # ---- find_order, timed separately (not part of the metric)
    eng.rewind()
    fo_ms0 = eng.stats()["ms_find_order"]
    for first, cnt in sched:
        eng.divide_rounds(first, cnt)
        eng.find_order(eng.decide_fame())
    fo_ms = eng.stats()["ms_find_order"] - fo_ms0
    n_ordered = eng.n_transactions

    # ---- end-to-end leg from host buffers
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    e2e_ms = 0.0
    e2e_steps = args.steps
    for _ in range(e2e_steps):
        e2e_ms += step_e2e()
    st_e2e = eng.stats()              # sw_reset clears the counters: this is the last e2e step alone
    barrier()

    dev_ms_max, e2e_ms_max, wall_ms_max = max_over_ranks([dev_ms, e2e_ms, wall_ms], dist, "cuda")

    if rank == 0:
        peak, peak_src = load_peaks()
        launches = timed_launches[0]
        ms_div = st1["ms_divide_rounds"] - st0["ms_divide_rounds"]
        ms_fame = st1["ms_decide_fame"] - st0["ms_decide_fame"]
        n_div_launch = len(sched) * args.steps
        ms_cs = st1["ms_can_see"] - st0["ms_can_see"]
        ms_rk = st1["ms_rounds_kernel"] - st0["ms_rounds_kernel"]
        impl = os.environ.get("SW_DIVIDE_IMPL", "5")
        batch = impl not in ("3", "4")
        bpe = rounds_bytes_per_event(M) if batch else algorithmic_bytes_per_event(M)
        ms_dom = ms_rk if ms_rk > 0 else ms_div
        achieved = (N * args.steps * bpe) / (ms_dom * 1e-3) / 1e9
        cs_achieved = (N * args.steps * can_see_bytes_per_event(M)) / (ms_cs * 1e-3) / 1e9 if ms_cs > 0 else None
        path_achieved = (N * args.steps * algorithmic_bytes_per_event(M)) / ((ms_div + ms_fame) * 1e-3) / 1e9
        value = whole_job_rate(world, N, args.steps, dev_ms_max)
        e2e_value = whole_job_rate(world, N, e2e_steps, e2e_ms_max)
        # CPU baseline next to it: one pass of the oracle port on the box's host cores
        n_cpu, s_cpu, fo_cpu = run_cpu_pass(tr, K, limit=None if N <= 1000000 else 1000000)
        line = {
            "metric": "events/sec divide_rounds+decide_fame", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(wl, world),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "events/s",
                    # counted by the library from the copies it issued in the last step (event columns incl. the
                    # derived seq/height columns; results + per-call scalars)
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"]),
                    "ms_per_step": e2e_ms_max / e2e_steps,
                    "kernel_ms_last_step": {"divide_rounds": st_e2e["ms_divide_rounds"], "can_see_scan": st_e2e["ms_can_see"],
                                            "rounds_kernel": st_e2e["ms_rounds_kernel"], "decide_fame": st_e2e["ms_decide_fame"]},
                    "what": "reset + per chunk: sw_append (host checks, pinned host -> HBM) + sw_divide_rounds + "
                            "sw_decide_fame, appends issued two chunks ahead of the consensus calls so that the copies "
                            "and the can_see scan overlap the round kernels; then round/witness/famous of every event "
                            "back to pinned host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": ("k_rounds_batch (round numbers of a chunk: one cooperative launch per divide_rounds "
                                    "call, one grid-wide step per round)") if batch else
                                   "k_divide_levels (level walker: can_see rows + rounds + witnesses)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ((ROUNDS_DRAM_BYTES_PER_LAUNCH if batch else WALKER_DRAM_BYTES_PER_LAUNCH)
                                     if (M == 64 and K == 65536) else None),
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                           "(profiles/README.md)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_event": bpe, "events_per_launch": K,
                         "ms_per_launch": ms_dom / n_div_launch,
                         "note": ("latency-bound, not HBM-bound: the depth of the computation is the number of rounds "
                                  "(one grid barrier pair per round, ~1449 rounds per 1M events at M=64); cycle "
                                  "accounting in profiles/ and tools/rounds_cycles.py") if batch else
                                 "latency/issue-bound: one CTA walks ~53.7k dependent levels per 1M events at M=64"},
            "roofline_can_see": None if cs_achieved is None else {
                "bound": "hbm", "kernel": "k_cs_local<.,1> + k_cs_collect + k_cs_boundary + k_cs_local<.,2> (blocked max-plus scan)",
                "achieved": cs_achieved, "peak": peak, "unit": "GB/s", "frac": cs_achieved / peak,
                "algorithmic_bytes_per_event": can_see_bytes_per_event(M), "ms_per_step": ms_cs / args.steps,
                "note": "the resident leg scans all appended events in the first divide_rounds call of a step (the "
                        "end-to-end leg scans chunk by chunk, beside the round kernel of the previous chunk)"},
            "roofline_path": {"bound": "hbm", "what": "all kernels of divide_rounds + decide_fame, SURVEY.md 8d B(M)",
                              "achieved": path_achieved, "peak": peak, "unit": "GB/s", "frac": path_achieved / peak,
                              "algorithmic_bytes_per_event": algorithmic_bytes_per_event(M)},
            "kernel_ms_per_step": {"divide_rounds": ms_div / args.steps, "decide_fame": ms_fame / args.steps,
                                   "can_see_scan": ms_cs / args.steps, "rounds_kernel": ms_rk / args.steps,
                                   "wall": wall_ms_max / args.steps},
            "impl": {"divide": {"5": "5 (round batch)", "4": "4 (level walker)", "3": "3 (per-event flags)"}.get(impl, impl),
                     "can_see": "scan" if batch else os.environ.get("SW_CANSEE_IMPL", "fused")},
            "cpu_baseline": {"value": n_cpu / s_cpu, "unit": "events/s", "cores": 1, "kind": "port",
                             "sample": "full trace (%d events), oracle/swirld_oracle.c single thread on a host with %d cpus"
                                       % (n_cpu, os.cpu_count())},
            "find_order": {"events_per_s": n_ordered / (fo_ms * 1e-3) if fo_ms > 0 else None, "ordered": int(n_ordered),
                           "ms": fo_ms, "cpu_port_events_per_s": n_cpu / fo_cpu if fo_cpu > 0 else None},
            "checksum": {"round_sum": check[0], "max_round": check[1], "consensus_rounds": check[2]},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--events", type=int, default=0, help="override the workload's event count")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.events:
        wl["N"] = args.events
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        bench_reference(args, wl, rank, world)
    else:
        bench_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
