#!/usr/bin/env python
"""Benchmark of the stabilised G(tau=0) hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config C1|C2|C3|C4|C5] [--scaling weak|strong] [--impl reference]

A *step* is one pass of the hot path over one batch of chains: for every chain, L slices of
Hubbard-like B matrices, re-factorised every n_stab=10 slices (lmul! -> ldr!), then inv_IpA!
(inv_IpUV! for the complex config) -> G, log|det G|, sign.  `value` = evaluations/s with the inputs
already resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same
metric through the public host API with HOST (pinned) buffers, host<->device copies inside the timed
region (the device->host read of step k overlaps the computation of step k+1 on a copy stream, double-buffered
outputs).  Multi-GPU: the batch of chains is sharded -- `--scaling weak` (default): each rank runs the config's
per-GPU share of chains (C5: 512); `--scaling strong`: the config's TOTAL number of chains (C5: 4096) is split over
the ranks -- and a single small NCCL all-reduce per step combines the cross-walker sums of log|det| and sign.  `--impl reference` times the CPU restatement of the reference (oracle/, C over ILP64 OpenBLAS
LAPACK; Julia is not installable here) on all host cores, one chain per thread, 1 BLAS thread each.
The oracle is executed in exactly two places: that reference arm and the `cpu_baseline` leg (rank 0, one GPU); the
`parity` key compares the GPU results of chains 0..1 with the results the cpu_baseline leg computed for the same chains.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stabilised G(tau=0) evals/sec"
UNIT = "evals/s"


# chains per GPU under weak scaling where it differs from the config's total (BASELINE: "4096 chains ... sharded over
# 1/2/4/8 B200" = 512 per GPU at 8 GPUs); every other config's batch is its per-GPU share
WEAK_CHAINS_PER_GPU = {"C5": 512}
DEFAULT_STEPS = {"C1": 50, "C2": 20, "C3": 10, "C4": 3, "C5": 5}


def get_cfg(name, chains=None, scaling="weak", world=1, rank=0):
    """The workload of this rank: (config with .chains = chains of THIS rank, global chain ids, total chains)."""
    import sla_b200
    from sla_b200 import walkers
    syn = sla_b200.synthetic
    base = syn.CONFIGS[name]
    if scaling == "strong":
        total = chains if chains is not None else base.chains
        ids = walkers.shard_chains(total, world, rank)
    else:
        per = chains if chains is not None else WEAK_CHAINS_PER_GPU.get(name, base.chains)
        total = per * world
        ids = walkers.weak_chain_ids(per, rank)
    cfg = syn.HubbardConfig(base.name, base.Lx, base.L, base.n_stab, len(ids), dtype=base.dtype, U=base.U,
                            dtau=base.dtau, mu=base.mu, phase=base.phase, inverse=base.inverse)
    return cfg, ids, total


def config_json(cfg, n_gpus, extra=None, total=None, scaling="weak"):
    total = cfg.chains * n_gpus if total is None else total
    c = {"workload": f"{cfg.name}: {total} chains over {n_gpus} GPU(s) ({scaling} scaling, {cfg.chains} on rank 0), "
                     f"N={cfg.N} ({cfg.Lx}x{cfg.Lx} Hubbard, U={cfg.U}, "
                     f"dtau={cfg.dtau}), L={cfg.L}, n_stab={cfg.n_stab}, inv_{cfg.inverse}!",
         "N": cfg.N, "L": cfg.L, "n_stab": cfg.n_stab, "chains_per_gpu": cfg.chains,
         "chains_total": total, "eltype": "ComplexF64" if cfg.dtype.kind == "c" else "Float64",
         "inverse": "inv_" + cfg.inverse + "!", "parallelism": f"chains sharded over {n_gpus} GPU(s)",
         "flops_per_eval": cfg.flops_per_eval(),
         "ldr_pivoting": "panel-restricted (SLA_LDR_PIVOT=panel; not the ?geqp3 pivot sequence)"
                         if os.environ.get("SLA_LDR_PIVOT", "").startswith("p") else "geqp3 (greedy, as the reference)"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's C restatement on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_run(cfg, chains, threads, keep=0, blas_threads=1):
    """Time the CPU arm on chains 0..chains-1; with `keep` also return (G, logdet, sign) of the first `keep` chains."""
    from oracle import c_oracle
    expK = cfg.expK()
    fields = cfg.fields(range(chains))
    t0 = time.perf_counter()
    G, ld, sg = c_oracle.eval_batch(expK, fields, cfg.lam, cfg.n_stab, cfg.inverse, nthreads=threads,
                                    blas_threads=blas_threads)
    t = time.perf_counter() - t0
    return (t, (G[:keep], ld[:keep], sg[:keep])) if keep else t


def host_cores():
    """(logical CPUs this process may use, physical cores of the box)."""
    logical = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        physical = psutil.cpu_count(logical=False) or logical
    except Exception:
        physical = logical
    return logical, physical


def cpu_baseline(cfg, budget_s=12.0, keep=2):
    """The cpu_baseline leg.  Its first pass evaluates chains 0..cores-1 -- the ids rank 0 evaluates on the GPU -- so the
    results of the first `keep` chains double as the parity check of the line (no second run of the oracle)."""
    from oracle import c_oracle
    cores, physical = host_cores()
    keep = min(keep, cores)
    t1, ref = cpu_run(cfg, cores, cores, keep=keep)      # one chain per core
    rounds = max(1, min(8, int(budget_s / max(t1, 1e-3))))
    chains = cores * rounds
    t = cpu_run(cfg, chains, cores) if rounds > 1 else t1
    # for transparency: the other arrangement, ONE worker x all BLAS threads, on a short sample
    n2 = 2 if t1 < 4.0 else 1
    t2 = cpu_run(cfg, n2, 1, blas_threads=cores)
    out = {"value": chains / t, "unit": UNIT, "cores": cores, "physical_cores": physical, "kind": "port",
           "sample": f"{chains} chains of the same workload ({cfg.name} size), {cores} worker threads x 1 BLAS "
                     f"thread, {t:.2f} s wall; C restatement of the reference over {c_oracle.blas_config()}",
           "secondary": {"value": n2 / t2, "unit": UNIT, "arrangement": f"1 worker x {cores} BLAS threads",
                         "sample": f"{n2} chain(s), {t2:.2f} s wall"}}
    if cfg.chains == 1:
        # a single-chain config measures LATENCY on the GPU (one chain in flight); the like-for-like CPU figures are the
        # single-chain ones: one worker with one BLAS thread, and `secondary` (one worker, all BLAS threads)
        t3 = cpu_run(cfg, 4, 1)
        out["single_chain"] = {"value": 4 / t3, "unit": UNIT, "arrangement": "1 worker x 1 BLAS thread",
                               "sample": f"4 chains one after the other, {t3:.3f} s wall"}
        out["note"] = (f"{cfg.name} is ONE chain: `value` of the GPU line is 1 / latency; this cpu_baseline value runs {cores} "
                       "independent chains side by side (throughput). Single-chain CPU figures: `single_chain`, `secondary`.")
    return out, ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, _, total = get_cfg(args.config, args.chains, args.scaling, max(args.gpus, 1), 0)
    cores, physical = host_cores()
    from oracle import c_oracle
    chains = cores  # bounded sample per step: one chain per host thread
    if args.steps is None:
        args.steps = {"C1": 20, "C2": 5, "C3": 2, "C4": 1, "C5": 2}.get(cfg.name, 2)
    for _ in range(args.warmup):
        cpu_run(cfg, chains, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(cfg, chains, cores)
    t = time.perf_counter() - t0
    value = chains * args.steps / t
    sample = (f"{chains} chains/step ({cfg.name} size), {cores} worker threads ({physical} physical cores) x 1 BLAS thread; "
              f"LAPACK-level C restatement (Julia not installed); {c_oracle.blas_config()}")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "c128" if cfg.dtype.kind == "c" else "f64", "data": "synthetic",
           "config": config_json(cfg, args.gpus, None, total, args.scaling),
           "cpu_sample_chains_per_step": chains,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "physical_cores": physical, "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measure_dgemm_peak(torch, seconds=0.6):
    """Library FP64 GEMM throughput on this box (MEASURED_PEAKS.json has no FP64 figure)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b, c
    return best


def ncu_record(kernel, cfg_name):
    """ncu-derived figures of a kernel (DRAM bytes per launch, tensor-pipe activity) from the COMMITTED summary
    profiles/ncu_kernel_metrics.json, keyed by kernel and config; None when no capture of this build is recorded --
    they are never constants in this file."""
    path = os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")
    try:
        with open(path) as f:
            for e in json.load(f)["entries"]:
                if e["kernel"] == kernel and e["config"] == cfg_name:
                    return e
    except (OSError, ValueError, KeyError):
        pass
    return None


def measure_gemm_kernel(torch, sla, cfg, reps=20):
    """Throughput of the dominant kernel -- the batched DMMA GEMM that builds the group products (shared expK x
    per-chain matrix, row-scaled epilogue) -- in the pattern the step issues it: sla_greens_chain builds the group
    products one group per launch (batch = chains), round-robin over TWO low-priority side streams, so two launches are
    in flight together.  Timed with CUDA events on a parent stream that forks into / joins the two launching streams;
    returns (ms per concurrent pair, FLOPs per pair, matrices per pair, TFLOP/s of ONE launch running alone)."""
    n = cfg.N
    batch = cfg.chains
    dt = torch.complex128 if cfg.dtype.kind == "c" else torch.float64
    A = sla.to_device(cfg.expK())
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = []
    for st in streams:
        B = torch.randn(batch, n, n, dtype=torch.float64, device="cuda").to(dt)
        bufs.append([B, torch.empty_like(B)])
    torch.cuda.synchronize()
    flops1 = (8.0 if cfg.dtype.kind == "c" else 2.0) * n ** 3 * batch

    def run(nstreams, count):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event() for _ in range(nstreams)]
        e0.record(main)
        for i in range(nstreams):
            streams[i].wait_event(e0)
            with torch.cuda.stream(streams[i]):
                for _ in range(count):
                    sla.gemm(bufs[i][1], A, bufs[i][0])
                    bufs[i].reverse()
                ends[i].record(streams[i])
        for ev in ends:
            main.wait_event(ev)
        e1.record(main)
        e1.synchronize()
        return e0.elapsed_time(e1) / count

    run(2, 2); run(1, 2)                      # warm-up (also creates the per-stream library handles)
    ms_pair = run(2, reps)
    ms_single = run(1, reps)
    return ms_pair, 2 * flops1, 2 * batch, flops1 / (ms_single * 1e-3) / 1e12


def measure_ldr_kernel(torch, sla, cfg, F, reps=10):
    """Average duration of ONE ldr! (pivoted QR + R extraction + Q formation) of the whole batch on the kind of matrix
    the chain hands it -- (L D) of the final factorisation: columns graded over the full range of d -- with CUDA
    events; algorithmic work 8/3 N^3 per matrix (4/3 ?geqp3 + 4/3 ?orgqr), x4 complex."""
    A = (F.L * F.d[:, :, None].to(F.L.dtype)).contiguous()     # column j of L scaled by d_j (column-major batch)
    ws = sla.ldr_workspace(F)
    W = sla.ldr(F)
    for _ in range(2):
        sla.ldr_(W, A, ws)
    torch.cuda.synchronize()
    l0 = sla.load().sla_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sla.ldr_(W, A, ws)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = (sla.load().sla_launch_count() - l0) // reps
    flops = (4.0 if cfg.dtype.kind == "c" else 1.0) * (8.0 / 3.0) * cfg.N ** 3 * cfg.chains
    return ms, flops, sla.last_ldr_algo(), int(launches)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import sla_b200 as sla

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")

    torch.cuda.set_device(local)
    # everything runs on an explicit (non-legacy) stream: the fused driver replays its launches as a CUDA graph there
    torch.cuda.set_stream(torch.cuda.Stream())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(sla.LIB_PATH):      # fresh checkout without built artefacts: compile first (local rank 0)
        if local == 0:
            import __graft_entry__ as _g
            _g.build()
    if world > 1:
        dist.barrier()
    cfg, chain_ids, total_chains = get_cfg(args.config, args.chains, args.scaling, world, rank)
    n, L, chains = cfg.N, cfg.L, cfg.chains
    dt = torch.complex128 if cfg.dtype.kind == "c" else torch.float64
    from sla_b200 import walkers

    expK_h = torch.from_numpy(np.ascontiguousarray(cfg.expK().T)).pin_memory()      # column-major storage
    fields_h = torch.from_numpy(cfg.fields(chain_ids)).pin_memory()
    expK_d = expK_h.cuda()[None]
    fields_d = fields_h.cuda()
    # outputs are double-buffered: the device->host read of step k runs on a copy stream while step k+1 computes
    Gs = [torch.empty((chains, n, n), dtype=dt, device="cuda") for _ in range(2)]
    logdets = [torch.empty(chains, dtype=torch.float64, device="cuda") for _ in range(2)]
    signs = [torch.empty(chains, dtype=dt, device="cuda") for _ in range(2)]
    accs = [torch.zeros(4, dtype=torch.float64, device="cuda") for _ in range(2)]
    expK_in = [torch.empty_like(expK_d) for _ in range(2)]
    fields_in = [torch.empty_like(fields_d) for _ in range(2)]
    G, logdet, sign, acc = Gs[0], logdets[0], signs[0], accs[0]
    ws = torch.empty(sla.greens_chain_workspace_bytes(dt, n, chains, L, cfg.n_stab), dtype=torch.uint8, device="cuda")
    G_hs = [torch.empty((chains, n, n), dtype=dt).pin_memory() for _ in range(2)]
    logdet_hs = [torch.empty(chains, dtype=torch.float64).pin_memory() for _ in range(2)]
    sign_hs = [torch.empty(chains, dtype=dt).pin_memory() for _ in range(2)]
    acc_hs = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(2)]
    lib = sla.load()
    copy_stream = torch.cuda.Stream()
    ev_done = [torch.cuda.Event() for _ in range(2)]     # step's results complete on the compute stream
    ev_read = [torch.cuda.Event() for _ in range(2)]     # step's results are in host memory
    state = {"k": 0}

    def step_device():
        sla.greens_chain(expK_d, fields_d, cfg.lam, cfg.n_stab, cfg.inverse, G=G, logdet=logdet, sign=sign, ws=ws)
        # cross-walker reduction of log-determinants and signs: one library kernel + the only collective of the path
        walkers.reduce_walkers(logdet, sign, out=acc)

    ev_in = [torch.cuda.Event() for _ in range(2)]       # step's inputs have arrived in their device staging buffers

    def upload(i):
        """Host -> device copy of one step's inputs into staging set i, on the copy stream (after the chain that last
        read set i)."""
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[i])
            expK_in[i].copy_(expK_h[None], non_blocking=True)
            fields_in[i].copy_(fields_h, non_blocking=True)
            ev_in[i].record(copy_stream)

    def step_e2e():
        """One end-to-end step, software-pipelined over two buffer sets: while step k computes, the copy stream brings
        in the inputs of step k+1 and reads out the results of step k-1; every step's H2D and D2H are inside the timed
        region (the inputs of the first timed step are uploaded by the last warm-up step, the last step uploads a set
        that is not consumed: the count per step is the same)."""
        i = state["k"] & 1
        if state["k"] == 0:
            upload(0)
        state["k"] += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_read[i])                 # buffer set i was read out two steps ago
        cur.wait_event(ev_in[i])                   # this step's inputs are on the device
        sla.greens_chain(expK_in[i], fields_in[i], cfg.lam, cfg.n_stab, cfg.inverse, G=Gs[i], logdet=logdets[i],
                         sign=signs[i], ws=ws)
        walkers.reduce_walkers(logdets[i], signs[i], out=accs[i])
        ev_done[i].record(cur)
        upload(i ^ 1)                              # next step's inputs (its staging set was last read by step k-1)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[i])
            G_hs[i].copy_(Gs[i], non_blocking=True)
            logdet_hs[i].copy_(logdets[i], non_blocking=True)
            sign_hs[i].copy_(signs[i], non_blocking=True)
            acc_hs[i].copy_(accs[i], non_blocking=True)
            ev_read[i].record(copy_stream)

    def e2e_drain():
        copy_stream.synchronize()
        torch.cuda.current_stream().wait_stream(copy_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()                  # every step's results are in host memory before the clock stops
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps = args.steps if args.steps is not None else DEFAULT_STEPS.get(cfg.name, 10)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    l0 = lib.sla_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_device, steps)
    clocks = sampler.stop() if sampler else None
    launches = (lib.sla_launch_count() - l0) // steps
    value = total_chains * steps / (ms_total * 1e-3)
    ldr_algo = sla.last_ldr_algo()

    for _ in range(2):
        step_e2e()
    e2e_drain()
    ms_e2e = timed(step_e2e, steps, drain=e2e_drain)
    e2e_value = total_chains * steps / (ms_e2e * 1e-3)
    h2d = expK_h.numel() * expK_h.element_size() + fields_h.numel()
    d2h = sum(t.numel() * t.element_size() for t in (G_hs[0], logdet_hs[0], sign_hs[0], acc_hs[0]))
    last = (state["k"] - 1) & 1
    G_h, logdet_h, sign_h = G_hs[last], logdet_hs[last], sign_hs[last]

    parity = None   # filled from the cpu_baseline leg below (rank 0, one GPU): same chain ids on both sides

    roof = roof_ldr = cpub = None
    if rank == 0:
        gm_ms, gm_flops, gm_batch, gm_single = measure_gemm_kernel(torch, sla, cfg)
        peak = measure_dgemm_peak(torch)
        achieved = gm_flops / (gm_ms * 1e-3) / 1e12
        step_tflops = cfg.flops_per_eval() * chains / (ms_total / steps * 1e-3) / 1e12
        rec = ncu_record("gemm_kernel", cfg.name)
        roof = {"bound": "tensor", "kernel": "sla::gemm_kernel (batched FP64 DMMA GEMM, group-product shape)",
                "pattern": f"as sla_greens_chain issues it: one group per launch (batch {gm_batch // 2}), two launches in flight "
                           f"on two side streams; launch_ms / launch_batch / flops_per_launch describe one concurrent PAIR; "
                           f"a single launch of batch {gm_batch // 2} running alone: {gm_single:.2f} TFLOP/s",
                "single_launch_alone_tflops": gm_single,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": rec["dram_bytes_per_launch"] if rec else None,
                "traffic_source": (f"{rec['source']} @ {rec['commit']} (ncu dram read+write per launch)" if rec else None),
                "dmma_pipe_active_pct_ncu": rec.get("dmma_pipe_pct") if rec else None,
                "algorithmic_bytes_per_launch": (gm_batch * 2 + 1) * n * n * (16 if cfg.dtype.kind == "c" else 8),
                "peak_source": "FP64 cuBLAS DGEMM 8192^3 via torch.matmul measured in this run (MEASURED_PEAKS.json "
                               "holds only bf16 and HBM figures; nominal B200 FP64 ~37-40 TFLOP/s)",
                "launch_ms": gm_ms, "launch_batch": gm_batch, "flops_per_launch": gm_flops,
                "step_tflops": step_tflops, "step_frac_of_peak": step_tflops / peak}
        # second entry: the kernel furthest from its bound -- the LDR factorisation (pivoted QR + Q formation)
        F = sla.ldr(torch.empty((chains, n, n), dtype=dt, device="cuda"))
        if cfg.inverse == "IpA":
            sla.greens_chain(expK_d, fields_d, cfg.lam, cfg.n_stab, "IpA", G=G, logdet=logdet, sign=sign, out_F=F, ws=ws)
        else:   # the IpUV driver does not return its two factorisations: take the IpA chain's
            sla.greens_chain(expK_d, fields_d, cfg.lam, cfg.n_stab, "IpA", G=G, logdet=logdet, sign=sign, out_F=F, ws=ws)
        lm_ms, lm_flops, lm_algo, lm_launches = measure_ldr_kernel(torch, sla, cfg, F)
        del F
        lrec = ncu_record("ldr", cfg.name)
        l_ach = lm_flops / (lm_ms * 1e-3) / 1e12
        roof_ldr = {"bound": "tensor", "kernel": f"ldr! of the whole batch (LDR kernel family: {lm_algo}, {lm_launches} launch(es))",
                    "achieved": l_ach, "peak": peak, "unit": "TFLOP/s", "frac": l_ach / peak,
                    "traffic": lrec["dram_bytes_per_launch"] if lrec else None,
                    "traffic_source": (f"{lrec['source']} @ {lrec['commit']}" if lrec else None),
                    "launch_ms": lm_ms, "launch_batch": chains, "flops_per_launch": lm_flops,
                    "note": "algorithmic work 8/3 N^3 per matrix (4/3 pivoted QR + 4/3 Q formation); latency-bound "
                            "per-column dependency chain, see DESIGN.md section 4"}
        if world == 1 and not args.no_cpu:
            # restore the results of the timed configuration for the parity check below
            sla.greens_chain(expK_d, fields_d, cfg.lam, cfg.n_stab, cfg.inverse, G=G, logdet=logdet, sign=sign, ws=ws)
            torch.cuda.synchronize()
            cpub, (Gr, lr, sr) = cpu_baseline(cfg)
            k = min(len(lr), chains)
            if k > 0 and not args.no_parity:   # chains 0..k-1 of this rank against the same chains of the CPU leg
                Gg = sla.to_host(G[:k])
                ld_g, sg_g = logdet.cpu().numpy(), sign.cpu().numpy()
                parity = {"chains_checked": k, "ldr_kernel": ldr_algo,
                          "max_rel_err_G": float(max(np.max(np.abs(Gg[i] - Gr[i])) / np.max(np.abs(Gr[i])) for i in range(k))),
                          "max_abs_err_logdet": float(np.max(np.abs(ld_g[:k] - lr[:k]))),
                          "sign_mismatch": int(np.sum(np.abs(sg_g[:k] - sr[:k]) > 1e-9)),
                          "e2e_host_copy_matches_device": bool(np.array_equal(np.swapaxes(G_h[:k].numpy(), 1, 2), Gg))}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
               "warmup": warm, "ms_per_step": ms_total / steps, "higher_is_better": True,
               "scaling": args.scaling, "vs_baseline": None, "dtype": "c128" if cfg.dtype.kind == "c" else "f64",
               "data": "synthetic",
               "config": config_json(cfg, world, {"l2_policy": "working set per step (group products + state, "
                                                  f"{ws.numel() / 2**30:.2f} GiB) is larger than the 126 MB L2"},
                                     total_chains, args.scaling),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": ms_e2e / steps,
                       "pipelining": "two buffer sets: the copy stream uploads the inputs of step k+1 and reads out the results "
                                     "of step k-1 while step k computes; the clock stops after the last step's results "
                                     "are in host memory"},
               "gpu_launches": int(launches), "ldr_kernel": ldr_algo, "clocks": clocks, "roofline": roof,
               "roofline_ldr": roof_ldr, "cpu_baseline": cpub, "parity": parity}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_flops(cfg):
    """Algorithmic FLOPs of one stabilised sweep of one chain: group products 2N^3(L - L/n_stab),
    stack of left factorisations (2 + 8/3 + 2)N^3 per group, inv_IpA! 6N^3, wrapping 4N^3 per slice, and per group
    lmul! (2 + 8/3 + 2)N^3 + inv_IpUV! 38/3 N^3; x4 for complex."""
    c = 4.0 if cfg.dtype.kind == "c" else 1.0
    ng = cfg.L // cfg.n_stab
    return c * cfg.N ** 3 * (2.0 * (cfg.L - ng) + ng * (20.0 / 3.0) + 6.0 + 4.0 * cfg.L + ng * (20.0 / 3.0 + 38.0 / 3.0))


def run_sweep(args):
    """`--workload sweep`: the fused sweep stepper (sla_sweep) on one BASELINE config -- a secondary
    record, same JSON contract (value device-resident, e2e through host buffers, parity against oracle/sweep.py)."""
    import torch
    import torch.distributed as dist

    import sla_b200 as sla

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, chain_ids, total_chains = get_cfg(args.config, args.chains, args.scaling, world, rank)
    n, L, chains, ng = cfg.N, cfg.L, cfg.chains, cfg.L // cfg.n_stab
    dt = torch.complex128 if cfg.dtype.kind == "c" else torch.float64
    expK = cfg.expK()
    expKinv = np.linalg.inv(expK)
    expK_h = torch.from_numpy(np.ascontiguousarray(expK.T)).pin_memory()
    expKinv_h = torch.from_numpy(np.ascontiguousarray(expKinv.T)).pin_memory()
    fields_h = torch.from_numpy(cfg.fields(chain_ids)).pin_memory()
    expK_d, expKinv_d, fields_d = expK_h.cuda()[None], expKinv_h.cuda()[None], fields_h.cuda()
    out = sla.sweep(expK_d, expKinv_d, fields_d, cfg.lam, cfg.n_stab)
    bufs = {k: out[k] for k in ("G", "G0", "logdet", "sign", "dG")}
    ws = out["ws"]
    host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in bufs.items() if k != "G0"}
    lib = sla.load()

    def step_device():
        sla.sweep(expK_d, expKinv_d, fields_d, cfg.lam, cfg.n_stab, ws=ws, **bufs)

    def step_e2e():
        ek = expK_h.to("cuda", non_blocking=True)[None]
        eki = expKinv_h.to("cuda", non_blocking=True)[None]
        fd = fields_h.to("cuda", non_blocking=True)
        sla.sweep(ek, eki, fd, cfg.lam, cfg.n_stab, ws=ws, **bufs)
        for k, h in host.items():
            h.copy_(bufs[k], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps = args.steps if args.steps is not None else max(2, DEFAULT_STEPS.get(cfg.name, 10) // 4)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    l0 = lib.sla_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_device, steps)
    clocks = sampler.stop() if sampler else None
    launches = (lib.sla_launch_count() - l0) // steps
    step_e2e()
    ms_e2e = timed(step_e2e, steps)
    parity = cpub = roof = None
    if rank == 0:
        peak = measure_dgemm_peak(torch)
        tfl = sweep_flops(cfg) * chains / (ms_total / steps * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "whole sweep (GEMM-dominated: group products, wrapping, stabilisation)",
                "achieved": tfl, "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak, "traffic": None,
                "peak_source": "FP64 cuBLAS DGEMM 8192^3 via torch.matmul measured in this run"}
        if world == 1 and not args.no_cpu:
            from oracle import sweep as osweep
            t0 = time.perf_counter()
            ref = osweep.sweep(expK, expKinv, cfg.fields(chain_ids[:1])[0], cfg.lam, cfg.n_stab)
            dt_cpu = time.perf_counter() - t0
            cpub = {"value": 1.0 / dt_cpu, "unit": "sweeps/s", "cores": int(os.environ.get("OMP_NUM_THREADS", host_cores()[0])),
                    "kind": "port", "sample": f"1 chain of the same workload, {dt_cpu:.2f} s wall; numpy/LAPACK restatement "
                    "oracle/sweep.py (one process, the BLAS library's own threads)"}
            Gg = sla.to_host(bufs["G"][:1])[0]
            ldg, sgg = bufs["logdet"][:, 0].cpu().numpy(), bufs["sign"][:, 0].cpu().numpy()
            lref = np.concatenate([[ref["logdet0"]], ref["logdet"]])
            sref = np.concatenate([[ref["sign0"]], ref["sign"]])
            parity = {"chains_checked": 1, "max_rel_err_G": float(np.max(np.abs(Gg - ref["G"])) / np.max(np.abs(ref["G"]))),
                      "max_abs_err_logdet": float(np.max(np.abs(ldg - lref))), "sign_mismatch": int(np.sum(np.abs(sgg - sref) > 1e-9)),
                      "max_wrap_error_chain0_gpu": float(bufs["dG"][:, 0].max().item()), "max_wrap_error_chain0_cpu": float(ref["dG"].max()),
                      "max_wrap_error_all_chains_gpu": float(bufs["dG"].max().item())}
        h2d = 2 * expK_h.numel() * expK_h.element_size() + fields_h.numel()
        d2h = sum(t.numel() * t.element_size() for t in host.values())
        line = {"metric": "stabilised DQMC sweeps/sec (setup stack + L wraps + L/n_stab re-stabilisations per chain), "
                          "fields fixed; max rel err vs CPU", "value": total_chains * steps / (ms_total * 1e-3),
                "unit": "sweeps/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_total / steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "c128" if cfg.dtype.kind == "c" else "f64", "data": "synthetic",
                "config": dict(config_json(cfg, world, {"l2_policy": f"working set per step ({ws.numel() / 2**30:.2f} GiB) is larger than the 126 MB L2"},
                                           total_chains, args.scaling), workload=f"sweep stepper on {cfg.name}: {total_chains} chains, N={n}, L={L}, n_stab={cfg.n_stab}",
                               flops_per_sweep=sweep_flops(cfg)),
                "e2e": {"value": total_chains * steps / (ms_e2e * 1e-3), "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpub, "parity": parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: per config, C2 = 20)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's per-GPU share on every rank; strong: the config's total split over the ranks")
    ap.add_argument("--chains", type=int, default=None,
                    help="chains per GPU (weak) or in total (strong); default: the config's")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--ldr-pivoting", default="geqp3", choices=["geqp3", "panel"],
                    help="geqp3: the reference's greedy column pivoting (default, strict); panel: panel-restricted pivoting "
                         "(SLA_LDR_PIVOT=panel, SURVEY 8(f)-4) -- a different pivot sequence, reported as its own record")
    ap.add_argument("--workload", default="greens", choices=["greens", "sweep"],
                    help="greens: stabilised G(tau=0) evaluations (the BASELINE metric); sweep: the fused sweep stepper")
    args = ap.parse_args()
    if args.ldr_pivoting == "panel":
        os.environ["SLA_LDR_PIVOT"] = "panel"      # read once by the library at its first factorisation
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "sweep":
        run_sweep(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
