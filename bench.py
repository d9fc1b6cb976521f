#!/usr/bin/env python
"""Benchmark of the GotenNet interaction path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE configs[1] — QM9-shape synthetic batch of 1024
molecules per GPU (avg 18 atoms, cutoff 5 A, max 32 neighbours), n_atom_basis=256,
n_interactions=4, lmax=2, model flags of configs/model/gotennet.yaml, fp32.
A step = radius graph + geometry + forward + backward (all parameter gradients) of
loss = h.sum() + X.pow(2).sum(); for N>1 one NCCL all-reduce of the flat fp32 gradient
buffer is inside the step.  value = molecules/s over all ranks, device-timed with CUDA
events (max over ranks), inputs resident in HBM.  e2e = the same through
GotenNetWrapper.forward with HOST (pinned) inputs: H2D of z/pos/batch and a D2H read of
the loss inside the timed region.

--impl reference times the CPU oracle (a PyTorch restatement of the reference, pinned to
golden vectors produced by the verbatim reference; the reference tree itself is not on
the GPU box) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = dict(n_atom_basis=256, n_interactions=4, lmax=2, n_rbf=32, num_heads=8, sep_htr=True, sep_dir=True,
             sep_tensor=True, scale_edge=False, edge_updates=True)
CUTOFF, MAX_NBR = 5.0, 32
METRIC = "molecules/sec (fwd+bwd) QM9-shape batch"
UNIT = "molecules/s"


# ----------------------------------------------------------------- helpers ----
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes_per_molecule(N, E, n_mol):
    """fp32 forward bytes per layer, x3 for forward+backward."""
    C, L, S, lmax, n_layers = MODEL["n_atom_basis"], 8, 5, 2, MODEL["n_interactions"]
    total = 0
    for i in range(n_layers):
        last = i == n_layers - 1
        total += 4 * (2 * N * (1 + L) * C + E * C * (1 if last else 2) + E * (L + 1)) + 16 * E \
            + 4 * C * C * (10 + 3 * S + (0 if last else 2 + lmax))
    return 3 * total / n_mol


def gemm_flops(M, N, K):
    return 2.0 * M * N * K


# -------------------------------------------------------------- our arm -------
def run_ours(args):
    import gotennet_b200 as g
    from gotennet_b200 import ops
    from gotennet_b200._lib import lib
    from gotennet_b200.synthetic import synth_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in the product)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = lib()

    torch.manual_seed(0)
    model = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(CUTOFF), max_num_neighbors=MAX_NBR, activation="swish",
                              **MODEL).to(dev)
    from gotennet_b200.parallel import FlatGradBuffer
    params = [p for p in model.parameters()]
    fbuf = FlatGradBuffer(params)  # one flat fp32 buffer -> one NCCL all-reduce per step

    B = args.batch
    z, pos, batch = synth_batch("qm9", B, seed=1000 + rank)  # molecules shard by graph: each rank owns B
    zh, ph, bh = z.pin_memory(), pos.pin_memory(), batch.pin_memory()
    zd, pd, bd = z.to(dev), pos.to(dev), batch.to(dev)

    class Data:
        pass

    gemm_log = []  # (flops, start_evt, end_evt) of every goten_gemm launch while profiling is on
    orig_gemm = ops.gemm
    prof = {"on": False}

    def timed_gemm(A, lda, ta, Bm, ldb, tb, Cm, ldc, M, N, K, **kw):
        if not prof["on"]:
            return orig_gemm(A, lda, ta, Bm, ldb, tb, Cm, ldc, M, N, K, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_gemm(A, lda, ta, Bm, ldb, tb, Cm, ldc, M, N, K, **kw)
        e1.record()
        gemm_log.append((gemm_flops(M, N, K), e0, e1))

    ops.gemm = timed_gemm

    def step(host_inputs: bool):
        d = Data()
        if host_inputs:
            d.z, d.pos, d.batch = (zh.to(dev, non_blocking=True), ph.to(dev, non_blocking=True),
                                   bh.to(dev, non_blocking=True))
        else:
            d.z, d.pos, d.batch = zd, pd, bd
        for p in params:
            p.grad = None
        h, X = model(d)
        loss = h.sum() + X.pow(2).sum()
        loss.backward()
        if world > 1:
            fbuf.all_reduce()
        if host_inputs:
            return float(loss.item())  # D2H read of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_inputs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step(host_inputs)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step(False)
    plan = model.last_plan
    N_nodes, E = plan.N, plan.E

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = L.cdll.goten_launch_count()
    prof["on"] = True
    ms_total = timed(args.steps, False)
    prof["on"] = False
    launches = (L.cdll.goten_launch_count() - launches0) // max(args.steps, 1)
    for _ in range(2):
        step(True)
    ms_e2e = timed(args.steps, True)
    clocks = sampler.stop() if sampler else None

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    # dominant kernel: the GEMM family (edge / node projections and their gradients)
    g_ms = sum(a.elapsed_time(b) for _, a, b in gemm_log)
    g_fl = sum(f for f, _, _ in gemm_log)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    impl = os.environ.get("GOTEN_GEMM", "auto")
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else None
    roofline = {
        "kernel": "goten_gemm (all nn.Linear forward/backward GEMMs, fp32 result accuracy)",
        "bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
        "frac": (achieved / bf16_peak) if achieved else None, "traffic": None,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF",
        "share_of_step": g_ms / ms_total if ms_total > 0 else None,
        "launches_per_step": len(gemm_log) // max(args.steps, 1),
        "note": "fp32-accurate GEMM: 3xTF32 costs 6x the bf16 MMA time per FLOP, exact-fp32 SIMT peaks near 75 TF/s",
    }
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"QM9-shape synthetic, {B} molecules/GPU (N={N_nodes} atoms, E={E} edges on rank 0), "
                                   "cutoff 5A, max 32 nbrs, n_atom_basis=256 n_interactions=4 lmax=2 fwd+bwd "
                                   "(graph build + all parameter gradients)",
                       "parallelism": f"molecules sharded by graph over {world} GPU(s); one NCCL all-reduce of the "
                                      "flat fp32 gradient buffer" if world > 1 else "single GPU",
                       "l2": "per-step working set (~12 GB of saved activations) exceeds the 126 MB L2; no explicit flush",
                       "gemm_impl": impl},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(z.numel() * 8 + pos.numel() * 4 + batch.numel() * 8),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "algorithmic_mb_per_molecule": algorithmic_bytes_per_molecule(N_nodes, E, B) / 1e6,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(budget_s=20.0)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------ CPU baseline / reference arm
def _oracle_step(n_mol, seed=0):
    from oracle import gotennet_oracle as orc
    cfg = orc.OracleConfig(cutoff=CUTOFF, max_num_neighbors=MAX_NBR, **MODEL)
    z, pos, batch = orc.synth_batch("qm9", n_mol, seed=seed)
    sd = {k: v.clone().requires_grad_("radial_basis" not in k) for k, v in orc.make_state_dict(cfg, 0).items()}

    def run():
        for v in sd.values():
            v.grad = None
        h, X = orc.wrapper_forward(sd, cfg, z, pos, batch)
        (h.sum() + X.pow(2).sum()).backward()

    return run


def cpu_baseline(budget_s=20.0, n_mol=32):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = _oracle_step(n_mol)
    run()  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 and (not times or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": n_mol / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_mol} QM9-shape molecules, same model/loss, fwd+bwd incl. graph build, best of {len(times)} "
                      f"after 1 warm-up ({best:.2f} s per pass); oracle = PyTorch restatement of the reference"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_mol = 32
    run = _oracle_step(n_mol)
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        run()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    value = n_mol / dt
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{n_mol} QM9-shape molecules per step (bounded sample of the 1024-molecule batch), {steps} steps"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "QM9-shape synthetic, n_atom_basis=256 n_interactions=4 lmax=2 fwd+bwd; CPU oracle "
                               f"(PyTorch restatement of the reference) on {cores} host threads, {n_mol}-molecule sample per step"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="molecules per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
