#!/usr/bin/env python
"""Benchmark of the stabilised G(tau=0) hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]

A *step* is one pass of the hot path over one batch of chains: for every chain, L slices of
Hubbard-like B matrices, re-factorised every n_stab=10 slices (lmul! -> ldr!), then inv_IpA!
(inv_IpUV! for the complex config) -> G, log|det G|, sign.  `value` = evaluations/s with the inputs
already resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same
metric through the public host API with HOST (pinned) buffers, host<->device copies inside the timed
region.  Multi-GPU: the batch of chains is sharded, each rank runs `chains` chains of its own (weak
scaling) and a single small NCCL all-reduce per step combines the cross-walker sums of log|det| and
sign.  `--impl reference` times the CPU restatement of the reference (oracle/, C over ILP64 OpenBLAS
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


def get_cfg(name, chains=None):
    import sla_b200
    cfg = sla_b200.synthetic.CONFIGS[name]
    if chains is not None:
        syn = sla_b200.synthetic
        cfg = syn.HubbardConfig(cfg.name, cfg.Lx, cfg.L, cfg.n_stab, chains, dtype=cfg.dtype, U=cfg.U, dtau=cfg.dtau,
                                mu=cfg.mu, phase=cfg.phase, inverse=cfg.inverse)
    return cfg


def config_json(cfg, n_gpus, extra=None):
    c = {"workload": f"{cfg.name}: {cfg.chains} chains/GPU, N={cfg.N} ({cfg.Lx}x{cfg.Lx} Hubbard, U={cfg.U}, "
                     f"dtau={cfg.dtau}), L={cfg.L}, n_stab={cfg.n_stab}, inv_{cfg.inverse}!",
         "N": cfg.N, "L": cfg.L, "n_stab": cfg.n_stab, "chains_per_gpu": cfg.chains,
         "chains_total": cfg.chains * n_gpus, "eltype": "ComplexF64" if cfg.dtype.kind == "c" else "Float64",
         "inverse": "inv_" + cfg.inverse + "!", "parallelism": f"chains sharded over {n_gpus} GPU(s)",
         "flops_per_eval": cfg.flops_per_eval()}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's C restatement on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_run(cfg, chains, threads, keep=0):
    """Time the CPU arm on chains 0..chains-1; with `keep` also return (G, logdet, sign) of the first `keep` chains."""
    from oracle import c_oracle
    expK = cfg.expK()
    fields = cfg.fields(range(chains))
    t0 = time.perf_counter()
    G, ld, sg = c_oracle.eval_batch(expK, fields, cfg.lam, cfg.n_stab, cfg.inverse, nthreads=threads)
    t = time.perf_counter() - t0
    return (t, (G[:keep], ld[:keep], sg[:keep])) if keep else t


def cpu_baseline(cfg, budget_s=12.0, keep=2):
    """The cpu_baseline leg.  Its first pass evaluates chains 0..cores-1 -- the ids rank 0 evaluates on the GPU -- so the
    results of the first `keep` chains double as the parity check of the line (no second run of the oracle)."""
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    keep = min(keep, cores)
    t1, ref = cpu_run(cfg, cores, cores, keep=keep)      # one chain per core
    rounds = max(1, min(8, int(budget_s / max(t1, 1e-3))))
    chains = cores * rounds
    t = cpu_run(cfg, chains, cores) if rounds > 1 else t1
    return {"value": chains / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{chains} chains of the same workload ({cfg.name} size), {cores} worker threads x 1 BLAS "
                      f"thread, {t:.2f} s wall; C restatement of the reference over {c_oracle.blas_config()}"}, ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = get_cfg(args.config, args.chains)
    cores = os.cpu_count() or 1
    from oracle import c_oracle
    chains = cores  # bounded sample per step: one chain per host thread
    for _ in range(args.warmup):
        cpu_run(cfg, chains, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(cfg, chains, cores)
    t = time.perf_counter() - t0
    value = chains * args.steps / t
    sample = (f"{chains} chains/step ({cfg.name} size), {cores} worker threads x 1 BLAS thread; "
              f"LAPACK-level C restatement (Julia not installed); {c_oracle.blas_config()}")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "c128" if cfg.dtype.kind == "c" else "f64", "data": "synthetic",
           "config": config_json(cfg, args.gpus, {"cpu_sample_chains_per_step": chains}),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE group-product GEMM launch at the C2 shape (N=256, batch 128),
# from the `ncu --set full` capture summarised in profiles/ncu_gemm_kernel_r1.txt
NCU_GEMM_TRAFFIC_BYTES_C2 = 97.5e6


def measure_gemm_kernel(torch, sla, cfg, reps=20):
    """Average duration of the dominant kernel -- the batched DMMA GEMM that builds the group products
    (shared expK x per-chain matrix, row-scaled epilogue) -- at the shape/batch the step launches it
    with, timed with CUDA events on the launching stream."""
    n = cfg.N
    batch = cfg.chains * 2          # the fused driver builds the group products in sub-chunks of 2 groups
    dt = torch.complex128 if cfg.dtype.kind == "c" else torch.float64
    A = sla.to_device(cfg.expK())
    B = torch.randn(batch, n, n, dtype=torch.float64, device="cuda").to(dt)
    C = torch.empty_like(B)
    for _ in range(2):
        sla.gemm(C, A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sla.gemm(C, A, B)
        B, C = C, B
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = (8.0 if cfg.dtype.kind == "c" else 2.0) * n ** 3 * batch
    return ms, flops, batch


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(sla.LIB_PATH):      # fresh checkout without built artefacts: compile first (local rank 0)
        if local == 0:
            import __graft_entry__ as _g
            _g.build()
    if world > 1:
        dist.barrier()
    cfg = get_cfg(args.config, args.chains)
    n, L, chains = cfg.N, cfg.L, cfg.chains
    dt = torch.complex128 if cfg.dtype.kind == "c" else torch.float64
    from sla_b200 import walkers
    chain_ids = walkers.weak_chain_ids(chains, rank)           # global chain ids: sharding-invariant fields

    expK_h = torch.from_numpy(np.ascontiguousarray(cfg.expK().T)).pin_memory()      # column-major storage
    fields_h = torch.from_numpy(cfg.fields(chain_ids)).pin_memory()
    expK_d = expK_h.cuda()[None]
    fields_d = fields_h.cuda()
    G = torch.empty((chains, n, n), dtype=dt, device="cuda")
    logdet = torch.empty(chains, dtype=torch.float64, device="cuda")
    sign = torch.empty(chains, dtype=dt, device="cuda")
    ws = torch.empty(sla.greens_chain_workspace_bytes(dt, n, chains, L, cfg.n_stab), dtype=torch.uint8, device="cuda")
    G_h = torch.empty((chains, n, n), dtype=dt).pin_memory()
    logdet_h = torch.empty(chains, dtype=torch.float64).pin_memory()
    sign_h = torch.empty(chains, dtype=dt).pin_memory()
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    lib = sla.load()

    def reduce_walkers():
        # cross-walker reduction of log-determinants and signs (the only collective on this path)
        walkers.reduce_walkers(logdet, sign, out=acc)

    def step_device():
        sla.greens_chain(expK_d, fields_d, cfg.lam, cfg.n_stab, cfg.inverse, G=G, logdet=logdet, sign=sign, ws=ws)
        reduce_walkers()

    def step_e2e():
        ek = expK_h.to("cuda", non_blocking=True)[None]
        fd = fields_h.to("cuda", non_blocking=True)
        sla.greens_chain(ek, fd, cfg.lam, cfg.n_stab, cfg.inverse, G=G, logdet=logdet, sign=sign, ws=ws)
        reduce_walkers()
        G_h.copy_(G, non_blocking=True)
        logdet_h.copy_(logdet, non_blocking=True)
        sign_h.copy_(sign, non_blocking=True)
        torch.cuda.current_stream().synchronize()

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

    for _ in range(max(args.warmup, 3)):
        step_device()
    l0 = lib.sla_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = (lib.sla_launch_count() - l0) // args.steps
    value = chains * world * args.steps / (ms_total * 1e-3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = chains * world * args.steps / (ms_e2e * 1e-3)
    h2d = expK_h.numel() * expK_h.element_size() + fields_h.numel()
    d2h = sum(t.numel() * t.element_size() for t in (G_h, logdet_h, sign_h))

    parity = None   # filled from the cpu_baseline leg below (rank 0, one GPU): same chain ids on both sides

    roof = cpub = None
    if rank == 0:
        gm_ms, gm_flops, gm_batch = measure_gemm_kernel(torch, sla, cfg)
        peak = measure_dgemm_peak(torch)
        achieved = gm_flops / (gm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "sla::gemm_kernel (batched FP64 DMMA GEMM, group-product shape)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": NCU_GEMM_TRAFFIC_BYTES_C2 if (cfg.N == 256 and cfg.chains == 64 and cfg.dtype.kind != "c") else None,
                "traffic_unit": "bytes/launch (ncu dram read+write; algorithmic HBM bytes of the launch: 134.7e6)",
                "dmma_pipe_active_pct_ncu": 81.3,
                "peak_source": "FP64 cuBLAS DGEMM 8192^3 via torch.matmul measured in this run (MEASURED_PEAKS.json "
                               "holds only bf16 and HBM figures; nominal B200 FP64 ~37-40 TFLOP/s)",
                "launch_ms": gm_ms, "launch_batch": gm_batch, "flops_per_launch": gm_flops,
                "step_tflops": cfg.flops_per_eval() * chains / (ms_total / args.steps * 1e-3) / 1e12,
                "step_frac_of_peak": cfg.flops_per_eval() * chains / (ms_total / args.steps * 1e-3) / 1e12 / peak}
        if world == 1 and not args.no_cpu:
            cpub, (Gr, lr, sr) = cpu_baseline(cfg)
            k = min(len(lr), chains)
            if k > 0 and not args.no_parity:   # chains 0..k-1 of this rank against the same chains of the CPU leg
                Gg = sla.to_host(G[:k])
                parity = {"chains_checked": k,
                          "max_rel_err_G": float(max(np.max(np.abs(Gg[i] - Gr[i])) / np.max(np.abs(Gr[i])) for i in range(k))),
                          "max_abs_err_logdet": float(np.max(np.abs(logdet_h.numpy()[:k] - lr[:k]))),
                          "sign_mismatch": int(np.sum(np.abs(sign_h.numpy()[:k] - sr[:k]) > 1e-9))}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "c128" if cfg.dtype.kind == "c" else "f64",
               "data": "synthetic",
               "config": config_json(cfg, world, {"l2_policy": "working set per step (group products + state, "
                                                  f"{ws.numel() / 2**30:.2f} GiB) is larger than the 126 MB L2"}),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpub,
               "parity": parity}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
