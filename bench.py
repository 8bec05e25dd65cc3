#!/usr/bin/env python
"""
bench.py -- headline benchmark: SA flip-evaluations/sec (all chains, device-timed) on the 96x96 NiPt lattice.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun by the driver)
  python bench.py --impl reference --steps K --warmup W    (the reference's CPU algorithm on the host cores)

One "step" = one sweep (N = 9216 Metropolis steps) of every chain.  Workload (BASELINE.json configs[3]/[4]):
d = 96, 65,536 chains per GPU, reference-width MLP 49 -> 20 -> 1 on the 7x7 local descriptor
(OML/dynamic_cat.py:325-341), random-init weights (sklearn Glorot bounds), gate off, geometric ('exp') cooling
from T_0 = 0.05, x0 ~ Bernoulli(0.5), Philox4x32-10 proposals/uniforms.  Chains shard across ranks with no
data-path collective; at N > 1 a best-structure all-gather (NCCL) runs after every sweep inside the timed region.
--gpus 8 runs BASELINE configs[4] by default (524,288 chains, parallel tempering: ladders of 8 fixed temperatures
geometric in [T0*e^-5, T0] strided across the GPUs, a swap round after every sweep = all-gather of the objectives + a
deterministic swap kernel, best-structure all-gather every 10 sweeps); --pt / --no-pt force either mode at any N.  After
the timed region the PT run checks itself (temperature multiset preserved, every rank holds the same temperature
table, one extra swap round equals oracle/exchange.py on sampled ladders) and prints the result next to the value.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))

import numpy as np

D, K_WIN, H, T0 = 96, 49, 20, 0.05
METRIC = "SA flip-evaluations/sec (all chains, device-timed)"
UNIT = "flip-evals/s"


def glorot_mlp(offsets, hidden, seed):
    """Random-init weights at the reference layer widths (sklearn's Glorot-uniform bounds)."""
    rng = np.random.default_rng(seed)
    k = len(offsets)
    b = np.sqrt(6.0 / (k + hidden))
    W1, b1 = rng.uniform(-b, b, (k, hidden)), rng.uniform(-b, b, hidden)
    b = np.sqrt(6.0 / (hidden + 1))
    return W1, b1, rng.uniform(-b, b, hidden), float(rng.uniform(-b, b))


def exp_schedule(T_0, total_steps):
    step = np.arange(total_steps, dtype=np.float64)
    return T_0 * np.exp(-step / (total_steps / 5.0))          # OML/optimize_SA.py:26,64


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([v.strip() for v in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """The reference's algorithm (oracle port of OML/optimize_SA.py:59-132) on one chain: O(N*K) descriptor-matrix
    copy + per-row update + full surrogate evaluation per step."""
    seed, d, n_steps, total_steps, step0 = args
    from oracle import lattice as L, surrogate as S, sa as SA, philox
    offsets = L.local_offsets(3)
    W1, b1, w2, b2 = glorot_mlp(offsets, H, 0)
    p = S.SurrogateParams(offsets, W1, b1, w2, b2)

    class Sur(object):
        def eval_rate(self, M):
            return S.eval_rate_fp64(p, M)
    N = d * d
    rng = np.random.default_rng(seed)
    x0 = (rng.random(N) < 0.5).astype(np.uint8)
    temps = exp_schedule(T0, total_steps)[step0:step0 + n_steps]
    prop, u = philox.draws(seed, seed, np.arange(step0, step0 + n_steps), N)
    t0 = time.perf_counter()
    SA.optimize_faithful(x0, d, Sur(), offsets, temps, prop, u, n_record=1)
    return time.perf_counter() - t0


def _one_thread_per_worker():
    """One chain per core, as the reference runs one structure per MPI rank: keep each worker's BLAS single-threaded
    (otherwise cores x cores threads oversubscribe the host and the baseline is unfairly slow)."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"


def cpu_reference_rate(d, flips_per_core, cores, total_steps):
    """flip-evals/s of the reference CPU path with one independent chain per host core
    (drivers/Online_learn_main.py:108-114 runs one structure per MPI rank)."""
    import multiprocessing as mp
    _one_thread_per_worker()
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, d, 2, total_steps, 0) for i in range(cores)])      # warm-up (imports, BLAS)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, d, flips_per_core, total_steps, 0) for i in range(cores)])
        wall = time.perf_counter() - t0
    return cores * flips_per_core / wall, wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    d = a.d
    total = (a.warmup + a.steps) * d * d
    flips = a.ref_flips
    import multiprocessing as mp
    _one_thread_per_worker()
    ctx = mp.get_context("spawn")
    per_step = []
    with ctx.Pool(cores) as pool:
        for s in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(i, d, flips, total, s * flips) for i in range(cores)])
            per_step.append(time.perf_counter() - t0)
    timed = per_step[a.warmup:]
    value = cores * flips * len(timed) / sum(timed)
    sample = "%d independent chains (one per core) x %d Metropolis steps per bench step, d=%d, K=49 local MLP" % (
        cores, flips, d)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sum(timed) / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a), "note": "reference CPU algorithm (oracle port of OML/optimize_SA.py:59-132), bounded sample",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def use_pt(a):
    """BASELINE configs[4] (parallel tempering) is the 8-GPU configuration; configs[3] (cooling) everything else."""
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    return bool(a.pt) if a.pt is not None else world == 8


def workload_config(a, cores_note=None):
    if use_pt(a):
        wl = ("BASELINE configs[4]: %dx%d periodic NiPt lattice, %d chains per GPU (x%d GPUs), 49->20->1 MLP surrogate "
              "(7x7 local descriptor, random-init, gate off), parallel tempering: ladders of %d fixed temperatures geometric "
              "in [T0*e^-5, T0], T0=%g, strided across the GPUs; swap round after every sweep, best-structure all-gather "
              "every %d sweeps" % (a.d, a.d, a.chains, a.gpus, a.ladder, T0, a.gather_every or 10))
    else:
        wl = ("BASELINE configs[3]: %dx%d periodic NiPt lattice, %d chains per GPU, 49->20->1 MLP surrogate "
              "(7x7 local descriptor, random-init, gate off), exp cooling T0=%g" % (a.d, a.d, a.chains, T0))
    cfg = {"workload": wl,
           "lattice": "%dx%d" % (a.d, a.d), "chains_per_gpu": a.chains, "descriptor": "local7x7", "hidden": H,
           "step": "one sweep = %d Metropolis steps per chain" % (a.sweep_steps if a.sweep_steps > 0 else a.d * a.d),
           "l2_policy": "working set (%.1f GB of per-chain state) >> 126 MB L2; no flush needed" % (
               a.chains * a.d * a.d * (64 if a.compact else H * 4) / 1e9),
           "rng": "Philox4x32-10, key=seed, counter=(step, chain)",
           "state": ("first-layer state packed to 24 bits per hidden unit, 64 B per site (scores within 1e-5 of fp64)"
                     if a.compact else "first-layer state int32, 80 B per site")}
    if cores_note:
        cfg["note"] = cores_note
    return cfg


# ---------------------------------------------------------------------------------------------------------------
def pt_self_check(pt, ch, tscale_initial, rounds_done, rank, world, dist, n_sample=64):
    """Untimed tail of a parallel-tempering run.  (1) the multiset of temperatures is the initial ladder's; (2) every
    rank holds the same temperature table and its chains carry their slice of it; (3) one EXTRA swap round equals
    oracle/exchange.py (the checker -- not timed, not shipped) on `n_sample` ladders spread over the whole table."""
    import torch
    n = ch.n
    ts = pt.tscale_all
    ok_multiset = bool(torch.equal(torch.sort(ts)[0], torch.sort(tscale_initial)[0]))
    # order-sensitive 64-bit checksum of the table, compared across ranks
    bits = ts.view(torch.int64)
    mult = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64) * 2654435761
    chk = (bits * mult).sum().reshape(1)
    if dist is not None:
        allchk = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        ok_same = all(int(c) == int(chk) for c in allchk)
    else:
        ok_same = True
    ok_local = bool(np.array_equal(ch.get_tscale(), ts[rank * n:(rank + 1) * n].cpu().numpy()))
    before = ts.cpu().numpy().copy()
    pt.round(rounds_done, t_ref=T0)                      # the extra round (all ranks take part in its all-gather)
    after = pt.tscale_all.cpu().numpy()
    of_all = pt.of_all.cpu().numpy()
    ok_oracle, n_swapped = None, None
    if rank == 0:
        from oracle import exchange as OX
        n_global = world * n
        n_lad = n_global // pt.ladder
        lads = np.unique(np.linspace(0, n_lad - 1, min(n_sample, n_lad)).astype(np.int64))
        want = OX.exchange_round(of_all, before, world, n, pt.ladder, rounds_done & 1, pt.seed, rounds_done, T0, ladders=lads)
        pos = np.concatenate([[OX.gathered_position(l * pt.ladder + i, world, n) for i in range(pt.ladder)] for l in lads])
        ok_oracle = bool(np.array_equal(want[pos], after[pos]))
        n_swapped = int((before[pos] != after[pos]).sum())
    if dist is not None:
        flags = torch.tensor([int(ok_multiset), int(ok_local)], device=ts.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        ok_multiset, ok_local = bool(flags[0]), bool(flags[1])
    return {"temperature_multiset_preserved": ok_multiset, "tscale_table_identical_on_all_ranks": ok_same,
            "local_chains_carry_their_slice": ok_local, "extra_round_equals_oracle_exchange": ok_oracle,
            "sampled_ladders": int(min(n_sample, (world * n) // pt.ladder)), "temperatures_swapped_in_sample": n_swapped,
            "swap_rounds_run": int(rounds_done)}


def other_configs():
    """BASELINE configs[1] and configs[2] next to the headline (untimed tail of the default N=1 run; they are parity-test
    cases, the numbers are for the record): 4,096 chains on the reference's 12x12 lattice with the full 144-column
    descriptor and a tree gate (one sweep-equivalent of 14,400 steps, Philox draws), and batch scoring of 2^20 random
    structures (host bits in, host rates + LSC / NH3-exported histograms out)."""
    import torch
    from so_b200 import engine
    out = {}
    d, N = 12, 144
    lat = engine.Lattice(d, "NH3", device=torch.cuda.current_device())
    off = engine.full_offsets(d)
    W1, b1, w2, b2 = glorot_mlp(off, H, 0)
    rng = np.random.default_rng(0)
    depth = 8                                              # synthetic depth-8 gate on the 144 columns

    def node(level, t):
        i = len(t["feature"])
        for k, v in (("feature", -2), ("thr", 0.0), ("left", -1), ("right", -1), ("active", int(rng.random() < 0.6))):
            t[k].append(v)
        if level < depth:
            t["feature"][i], t["thr"][i] = int(rng.integers(0, N)), 0.5
            t["left"][i] = node(level + 1, t)
            t["right"][i] = node(level + 1, t)
        return i
    t = dict(feature=[], thr=[], left=[], right=[], active=[])
    node(0, t)
    tree = dict(feature=np.array(t["feature"], np.int32), thr=np.array(t["thr"]), left=np.array(t["left"], np.int32),
                right=np.array(t["right"], np.int32), active=np.array(t["active"], np.uint8))
    for name, tr in (("gated", tree), ("gate_off", None)):
        tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2, tree=tr)
        ch = engine.Chains(lat, 4096)
        ch.randomize(seed=1234, coverage=0.5)
        ch.attach(tab)
        temps = exp_schedule(T0, 2 * 14400)
        ch.anneal(temps[:14400], step0=0, seed=5)
        st = ch.anneal(temps[14400:], step0=14400, seed=5)
        out["configs[1] 12x12, 4096 chains x 14400 steps, full descriptor, " + name] = {
            "value": st["flips"] / (st["kernel_ms"] * 1e-3), "unit": UNIT, "kernel": st["kernel"], "ms": st["kernel_ms"],
            "accept_fraction": st["accepted"] / st["flips"]}
        ch.close()
    # the reference's surrogate is gated (OML/train_surrogate.py:34-54); the headline shape with a tree gate on the 7x7 window
    d96 = 96
    lat96 = engine.Lattice(d96, "NH3", device=torch.cuda.current_device())
    off7 = engine.local_offsets(3)
    W1l, b1l, w2l, b2l = glorot_mlp(off7, H, 0)
    rng7 = np.random.default_rng(0)

    def node7(level, t):
        i = len(t["feature"])
        for k, v in (("feature", -2), ("thr", 0.0), ("left", -1), ("right", -1), ("active", int(rng7.random() < 0.6))):
            t[k].append(v)
        if level < depth:
            t["feature"][i], t["thr"][i] = int(rng7.integers(0, 49)), 0.5
            t["left"][i] = node7(level + 1, t)
            t["right"][i] = node7(level + 1, t)
        return i
    t7 = dict(feature=[], thr=[], left=[], right=[], active=[])
    node7(0, t7)
    tree7 = dict(feature=np.array(t7["feature"], np.int32), thr=np.array(t7["thr"]), left=np.array(t7["left"], np.int32),
                 right=np.array(t7["right"], np.int32), active=np.array(t7["active"], np.uint8))
    tab = engine.SurrogateTables(lat96, off7, W1l, b1l, w2l, b2l, tree=tree7, compact=True)
    ch = engine.Chains(lat96, 16384)
    ch.randomize(seed=1234, coverage=0.5)
    ch.attach(tab)
    temps = exp_schedule(T0, 3 * d96 * d96)
    for i in range(3):
        st = ch.anneal(temps[i * d96 * d96:(i + 1) * d96 * d96], step0=i * d96 * d96, seed=5)
    out["96x96, 16384 chains, 7x7 window WITH a depth-8 tree gate, 24-bit packed state, third sweep"] = {
        "value": st["flips"] / (st["kernel_ms"] * 1e-3), "unit": UNIT, "kernel": st["kernel"], "ms": st["kernel_ms"],
        "accept_fraction": st["accepted"] / st["flips"]}
    ch.close()
    n = 1 << 20
    cov = rng.random((n, 1))
    bits = engine.pack_bits((rng.random((n, N)) < cov).astype(np.uint8))
    tab = engine.SurrogateTables(lat, off, W1, b1, w2, b2)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    bits_p = pin(bits.shape, torch.int32).view(np.uint32)
    bits_p[:] = bits
    res = (pin((n,), torch.float64), pin((n, 3), torch.int32), pin((n, 8), torch.int32))
    tab.score_batch(bits_p, out=res)
    t0 = time.perf_counter()
    for _ in range(3):
        tab.score_batch(bits_p, out=res)
    dt = (time.perf_counter() - t0) / 3
    out["configs[2] batch scoring, 2^20 structures, 12x12 (host in, host out)"] = {
        "value": n / dt, "unit": "structures/s", "ms": 1e3 * dt, "h2d_bytes": int(bits.nbytes),
        "d2h_bytes": int(sum(r.nbytes for r in res))}
    return out


# constant temperatures at which the bench workload (96x96, 7x7-window MLP, Glorot seed 0) accepts about 0.5 / 0.2 / 0.05
# of its proposals in equilibrium: profiles/r02_accept_calibration.txt (scripts/calibrate_accept.py)
ACCEPT_SWEEP_TEMPS = ((0.5, 1.55), (0.2, 0.87), (0.05, 0.565))


def accept_sweeps(ch, a, N, chain_id0, peak):
    """The hot kernel where flips ARE accepted: chains re-randomised, then for each calibrated constant temperature
    (descending) two sweeps of every chain to equilibrate and one measured sweep, same shape as the headline (which
    anneals from T0 = 0.05 and is frozen after its warm-up).  Returns [{target, T, accept_fraction (measured), ms, value, bytes_per_flip, achieved GB/s, frac}]."""
    out = []
    step0 = 10 ** 9                                  # Philox counters far away from the timed region
    ch.randomize(seed=4321 + chain_id0, coverage=0.5)     # the headline left the chains frozen: restart hot, cool stepwise
    for tgt, T in ACCEPT_SWEEP_TEMPS:
        for rep in range(3):                          # two sweeps to equilibrate at T, the third is the measured one
            st = ch.anneal(np.full(N, T), step0=step0, seed=a.seed + 1, chain_id0=chain_id0)
            step0 += N
        af = st["accepted"] / max(st["flips"], 1)
        b_alg = 4.0 * K_WIN * H * (1.0 + af) + 15.0 + 4.0 * af
        ach = st["flips"] * b_alg / (st["kernel_ms"] * 1e-3) / 1e9
        out.append({"target": tgt, "T": float(T), "accept_fraction": af, "ms": st["kernel_ms"],
                    "value": st["flips"] / (st["kernel_ms"] * 1e-3), "bytes_per_flip": b_alg, "achieved": ach,
                    "frac": ach / peak, "kernel": st["kernel"]})
    return out


# ---------------------------------------------------------------------------------------------------------------
def run_gpu(a):
    import torch
    from so_b200 import engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    d = a.d
    N = a.sweep_steps if a.sweep_steps > 0 else a.d * a.d          # Metropolis steps per chain per bench step
    n_local = a.chains
    lat = engine.Lattice(d, "NH3", device=local_rank)
    offsets = engine.local_offsets(3)
    W1, b1, w2, b2 = glorot_mlp(offsets, H, 0)
    tab = engine.SurrogateTables(lat, offsets, W1, b1, w2, b2, compact=bool(a.compact))
    ch = engine.Chains(lat, n_local)
    ch.randomize(seed=1234 + rank, coverage=0.5)
    ch.attach(tab)

    total_sweeps = a.warmup + a.steps
    temps_all = exp_schedule(T0, total_sweeps * N)
    chain_id0 = rank * n_local

    from so_b200 import parallel
    want_pt = use_pt(a)
    gather = parallel.BestGather(ch, chain_id0, world) if (world > 1 or want_pt) else None
    gather_every = a.gather_every or (10 if want_pt else 1)
    pt = None
    if want_pt:   # BASELINE configs[4]: fixed-temperature ladders (geometric in [T0*e^-5, T0]) + swap round per sweep
        pt = parallel.ReplicaExchange(ch, ladder=a.ladder, seed=a.seed, rank=rank, world=world)
        tscale_initial = pt.tscale_all.clone()
        temps_all = np.full(total_sweeps * N, T0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    n_gathers = [0]

    def sweep(s):
        st = ch.anneal(temps_all[s * N:(s + 1) * N], step0=s * N, seed=a.seed, chain_id0=chain_id0)
        coll_ms = 0.0
        if pt is not None or gather is not None:
            ev0.record()
            if pt is not None:
                pt.round(s, t_ref=T0)
            if gather is not None and ((s + 1) % gather_every == 0 or s == total_sweeps - 1):
                # every gather but the last overlaps the next sweep (side stream); the last one is waited for and timed here
                if pt is None and s != total_sweeps - 1 and s != a.warmup - 1:
                    gather.gather_async()
                else:
                    gather.finish()
                    gather.gather()
                n_gathers[0] += 1
            ev1.record()
            ev1.synchronize()
            coll_ms = ev0.elapsed_time(ev1)
        return st, coll_ms

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(a.warmup):
        sweep(s)
    barrier()
    n_gathers[0] = 0
    kern_ms, coll_ms, flips, accepted = 0.0, 0.0, 0, 0
    kernel_name = None
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for s in range(a.warmup, total_sweeps):
            st, cm = sweep(s)
            kern_ms += st["kernel_ms"]
            kernel_name = st["kernel"]                               # what the dispatch launched (so_anneal_stats.kernel_id)
            coll_ms += cm
            flips += st["flips"]
            accepted += st["accepted"]
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = kern_ms + coll_ms
    if dist is not None:
        t = torch.tensor([dev_ms, wall_ms, float(flips), float(accepted)], dtype=torch.float64, device="cuda")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms = float(mx[0]), float(mx[1])
        flips_all, accepted_all = float(sm[2]), float(sm[3])
    else:
        flips_all, accepted_all = float(flips), float(accepted)
    value = flips_all / (dev_ms * 1e-3)

    # end-to-end through the public API with HOST buffers: H2D occupancy -> state build -> sweep -> D2H result
    e2e = None
    if a.e2e_steps > 0:
        # pinned host buffers for the step's inputs and results
        bits_host = torch.empty((n_local, lat.W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        of = torch.empty(n_local, dtype=torch.float64).pin_memory().numpy()
        ch.get_bits(out=bits_host)
        barrier()
        t0 = time.perf_counter()
        for s in range(a.e2e_steps):
            ch.set_bits(bits_host)                                   # H2D + first-layer state rebuild
            ch.anneal(temps_all[-N:], step0=(total_sweeps + s) * N, seed=a.seed, chain_id0=chain_id0)
            ch.get_bits(out=bits_host)                               # D2H final occupancies
            ch.objective(out=of)                                     # D2H objective per chain
        barrier()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e = {"value": world * n_local * N * a.e2e_steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(bits_host.nbytes + N * 8), "d2h_bytes_per_step": int(bits_host.nbytes + of.nbytes),
               "steps": a.e2e_steps}

    pt_check = pt_self_check(pt, ch, tscale_initial, total_sweeps, rank, world, dist) if pt is not None else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    acc_frac = accepted_all / max(flips_all, 1.0)
    b_alg = 4.0 * K_WIN * H * (1.0 + acc_frac) + 15.0 + 4.0 * acc_frac       # SURVEY 8(d), bytes per flip-evaluation
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern_s = kern_ms * 1e-3
    achieved = (flips_all / world) * b_alg / kern_s / 1e9 if world == 1 else float(flips) * b_alg / kern_s / 1e9
    # DRAM bytes of one launch from the committed ncu capture of THIS command at THIS shape (profiles/); null otherwise
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            sh = tj.get("shape", {})
            if (sh.get("lattice") == "%dx%d" % (d, d) and sh.get("chains") == n_local and sh.get("steps_per_launch") == N
                    and str(kernel_name) in str(sh.get("kernel")) and bool(sh.get("compact", False)) == bool(a.compact)):
                traffic = tj.get("sa_window_kernel")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "wall_ms_per_step": wall_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("24-bit packed fixed-point state / int64 sums / f64 Metropolis" if a.compact else
                  "int32 fixed-point state / int64 sums / f64 Metropolis"),
        "data": "synthetic", "config": workload_config(a),
        "accept_fraction": acc_frac, "clocks": clk.summary(), "e2e": e2e,
        # SA kernel per sweep (+ arg-max kernel of the all-gather, + objective/decision/commit kernels of a swap round)
        "gpu_launches": a.steps * (1 + (3 if pt is not None else 0)) + n_gathers[0],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel_name,
                     "bytes_per_flip": b_alg, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                     "launch_ms": kern_ms / a.steps,
                     "note": "achieved / frac count SURVEY 8(d)'s algorithmic bytes (4*K*H*(1+a) + 15 + 4a: an int32 state); "
                             "the packed state moves fewer bytes than that, so frac can exceed 1 -- packed_state and dram "
                             "give the rate in the bytes of the layout in use and in measured DRAM bytes"},
    }
    if traffic:
        # measured DRAM bytes of one launch (ncu capture of this command, profiles/) over the launch time measured HERE
        line["roofline"]["dram"] = {"achieved": traffic / (kern_s / a.steps) / 1e9, "frac": traffic / (kern_s / a.steps) / 1e9 / peak,
                                    "unit": "GB/s", "what": "roofline.traffic / roofline.launch_ms: the DRAM the kernel actually "
                                                            "moves, as a fraction of the measured copy peak"}
    if a.compact:
        # the kernel actually moves 64 instead of 80 bytes per site: the same rate in bytes of the packed layout
        b_packed = 64.0 * K_WIN * (1.0 + acc_frac) + 15.0 + 4.0 * acc_frac
        line["roofline"]["packed_state"] = {"bytes_per_flip": b_packed, "achieved": achieved * b_packed / b_alg,
                                            "frac": achieved * b_packed / b_alg / peak,
                                            "what": "24-bit packed first-layer state, 64 B per site, records aligned with the "
                                                    "64-B L2 fill granule (DESIGN.md section 4)"}
    if gather is not None:
        line["collective_ms_per_step"] = coll_ms / a.steps
        line["best_structure"] = {"of": gather.best[0], "chain": gather.best[1]}
    if pt is not None:
        line["config"]["parallel_tempering"] = {"ladder": a.ladder, "swaps_local_rank0": pt.swapped}
        line["pt_self_check"] = pt_check
    if world == 1 and pt is None and a.accept_sweep:
        line["accept_sweep"] = accept_sweeps(ch, a, N, chain_id0, peak)
    if world == 1 and pt is None and a.compact and a.accept_sweep:
        # the same workload on the int32 state (80 B per site; the round-1 layout), same box, untimed tail
        tab32 = engine.SurrogateTables(lat, offsets, W1, b1, w2, b2, compact=False)
        ch.randomize(seed=1234 + rank, coverage=0.5)
        ch.attach(tab32)
        for s_ in range(a.warmup + 1):
            st32 = ch.anneal(temps_all[s_ * N:(s_ + 1) * N], step0=s_ * N, seed=a.seed, chain_id0=chain_id0)
        a32 = st32["accepted"] / max(st32["flips"], 1)
        b32 = 4.0 * K_WIN * H * (1.0 + a32) + 15.0 + 4.0 * a32
        v32 = st32["flips"] / (st32["kernel_ms"] * 1e-3)
        line["int32_state"] = {"value": v32, "unit": UNIT, "ms": st32["kernel_ms"], "accept_fraction": a32,
                               "frac": v32 * b32 / 1e9 / peak, "kernel": st32["kernel"],
                               "what": "sweep %d of the same schedule with the int32 first-layer state (--compact 0)" % a.warmup}
    if world == 1 and pt is None and a.other_configs:
        ch.close()
        line["other_configs"] = other_configs()
    if world == 1 and a.cpu_flips > 0:
        cores = os.cpu_count() or 1
        v, wall = cpu_reference_rate(d, a.cpu_flips, cores, total_sweeps * N)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d chains (one per core) x %d Metropolis steps of the reference algorithm "
                                          "(O(N*K) descriptor copy + full MLP evaluation per step), d=%d, %.1f s wall"
                                          % (cores, a.cpu_flips, d, wall)}
        # optimised CPU path, clearly NOT the reference: the oracle's C restatement of the delta algorithm
        # (O(K*H) per step, exact fixed point, OpenMP over chains) on the same lattice and surrogate
        try:
            os.environ["OMP_NUM_THREADS"] = str(cores)     # (the workers above were pinned to one BLAS thread each)
            from oracle import c_oracle, surrogate as OS, lattice as OL, philox as OP
            p_ = OS.SurrogateParams(OL.local_offsets(3), W1, b1, w2, b2)
            q_ = OS.quantize(p_)
            nc, ns = 4 * cores, 20000
            rng = np.random.default_rng(0)
            x0 = (rng.random((nc, d * d)) < 0.5).astype(np.uint8)
            prop, u = OP.draws(1, np.arange(nc)[:, None], np.arange(ns)[None, :], d * d)
            t0 = time.perf_counter()
            c_oracle.anneal_fixed_many(p_, q_, d, x0, exp_schedule(T0, ns), prop, u)
            dt = time.perf_counter() - t0
            line["cpu_delta_port"] = {"value": nc * ns / dt, "unit": UNIT, "cores": cores, "kind": "port (optimised, not the reference)",
                                      "sample": "%d chains x %d steps of the delta algorithm in C + OpenMP, incl. state build, %.1f s wall"
                                                % (nc, ns, dt)}
        except Exception as e:                       # the checker's C build is optional for the bench
            line["cpu_delta_port"] = {"unavailable": str(e)[:120]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_batch_score(a):
    """BASELINE configs[2]: batch scoring of random Ni/Pt occupancy structures (eval_init_structures analogue, no
    annealing) on one B200: surrogate rate + LSC and NH3-exported site-type histograms per structure."""
    import torch
    from so_b200 import engine
    d = a.score_d
    N = d * d
    n = a.structures
    rng = np.random.default_rng(2)
    cov = rng.random((n, 1))
    x = (rng.random((n, N)) < cov).astype(np.uint8)                 # randomize(coverage=None), dynamic_cat.py:73-76
    bits = engine.pack_bits(x)
    lat = engine.Lattice(d, "NH3", device=0)
    offsets = engine.full_offsets(d) if d <= 16 else engine.local_offsets(3)
    W1, b1, w2, b2 = glorot_mlp(offsets, H, 0)
    tab = engine.SurrogateTables(lat, offsets, W1, b1, w2, b2)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    bits_p = pin(bits.shape, torch.int32).view(np.uint32)
    bits_p[:] = bits
    out = (pin((n,), torch.float64), pin((n, 3), torch.int32), pin((n, 8), torch.int32))
    for _ in range(a.warmup):
        tab.score_batch(bits_p, out=out)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        rate, lh, ne = tab.score_batch(bits_p, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"metric": "structures scored/sec (host buffers in, host results out)", "value": n / dt,
                      "unit": "structures/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt,
                      "higher_is_better": True, "data": "synthetic", "dtype": "int32 fixed-point / int64 sums",
                      "config": {"workload": "BASELINE configs[2]: %d random structures, %dx%d, %s descriptor MLP %d->20->1 "
                                             "+ LSC / NH3-exported histograms" % (n, d, d, "full" if d <= 16 else "local7x7",
                                                                                  len(offsets)),
                                 "checksum_rate": float(rate.sum()), "checksum_hist": int(lh.sum() + ne.sum())},
                      "h2d_bytes_per_step": int(bits.nbytes), "d2h_bytes_per_step": int(rate.nbytes + lh.nbytes + ne.nbytes)}))


def run_train(a):
    """Regressor fit of train_surrogate.py:141-152 at the reference recipe's shape: 48 structures x 144
    sites x 3 rotations = 20,736 rows of the full 144-column descriptor; a step = one epoch (104 Adam steps of 200
    rows).  Device: so_trainer_* through DeviceMLPRegressor; beside it scikit-learn's own fit on the host cores."""
    import warnings
    from so_b200.train_surrogate import DeviceMLPRegressor
    warnings.simplefilter("ignore")
    n, K = 20736, 144
    rng = np.random.default_rng(0)
    X = (rng.random((n, K)) < 0.5).astype(np.float64)
    y = np.maximum(X @ (rng.normal(size=K) / np.sqrt(K)), 0) + 0.05 * rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    kw = dict(activation='relu', learning_rate_init=0.0001, alpha=0.1, hidden_layer_sizes=(20,), tol=1e-5, random_state=0)
    DeviceMLPRegressor(max_iter=a.warmup, **kw).fit(X, y)
    t0 = time.perf_counter()
    dev = DeviceMLPRegressor(max_iter=a.steps, **kw).fit(X, y)
    wall = time.perf_counter() - t0
    steps = -(-n // 200) * dev.n_iter_
    cpu = None
    if a.cpu_flips:
        from sklearn.neural_network import MLPRegressor
        t0 = time.perf_counter()
        ref = MLPRegressor(max_iter=a.steps, **kw).fit(X, y)
        t_ref = time.perf_counter() - t0
        cpu = {"value": steps / t_ref, "unit": "Adam steps/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": "sklearn.neural_network.MLPRegressor.fit, the call train_surrogate.py:150-152 makes, same data and "
                         "random_state, %d epochs, %.2f s wall" % (ref.n_iter_, t_ref),
               "loss_rel_diff": abs(ref.loss_curve_[-1] - dev.loss_curve_[-1]) / ref.loss_curve_[-1]}
    print(json.dumps({"metric": "Adam steps/sec of the regressor fit (200-row batches, device-timed)",
                      "value": steps / (1e-3 * dev.kernel_ms_), "unit": "Adam steps/s", "n_gpus": 1, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": dev.kernel_ms_ / dev.n_iter_, "higher_is_better": True,
                      "data": "synthetic", "dtype": "f64",
                      "config": {"workload": "regressor fit, %d rows x %d-column 0/1 descriptor, 144->20->1 ReLU MLP, Adam "
                                             "lr 1e-4, alpha 0.1, batch 200; one step = one epoch" % (n, K)},
                      "e2e": {"value": steps / wall, "unit": "Adam steps/s", "note": "DeviceMLPRegressor.fit wall time incl. "
                              "upload, host-generated epoch orders and read-back"},
                      "final_loss": dev.loss_curve_[-1], "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="anneal", choices=["anneal", "score", "train"])
    ap.add_argument("--structures", type=int, default=1 << 20)
    ap.add_argument("--score-d", type=int, default=12, help="lattice edge of the batch-scoring workload")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--d", type=int, default=D)
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU")
    ap.add_argument("--sweep-steps", type=int, default=0, help="Metropolis steps per chain per bench step (0 = d*d)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--pt", dest="pt", action="store_true", default=None,
                    help="parallel tempering (BASELINE configs[4]); default: on at --gpus 8, off otherwise")
    ap.add_argument("--no-pt", dest="pt", action="store_false")
    ap.add_argument("--gather-every", type=int, default=0, help="best-structure all-gather period in sweeps (0 = 10 with PT, 1 without)")
    ap.add_argument("--ladder", type=int, default=8, help="parallel-tempering ladder size")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--compact", type=int, default=1,
                    help="1 (default): 24-bit packed first-layer state, 64 B per site (so_surrogate_create_compact); 0: int32 state")
    ap.add_argument("--other-configs", type=int, default=1, help="1: also record BASELINE configs[1] and configs[2] numbers")
    ap.add_argument("--accept-sweep", type=int, default=1, help="1: also time one sweep at accept fractions 0.5/0.2/0.05")
    ap.add_argument("--cpu-flips", type=int, default=3000, help="Metropolis steps per core for the CPU baseline (0 = skip)")
    ap.add_argument("--ref-flips", type=int, default=1000, help="Metropolis steps per core per bench step, --impl reference")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "score":
        run_batch_score(a)
    elif a.workload == "train":
        run_train(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
