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

--impl reference times the CPU oracle (kind "port": a PyTorch restatement of the reference, pinned
to golden vectors produced by the verbatim reference; the reference tree itself cannot travel to
the GPU box) on all host cores; each of its --steps K steps is one 64-molecule micro-batch of the
same workload (the reference needs ~80 MB per molecule, so the 1024-molecule batch
is timed in micro-batches), after --warmup W untimed ones.

--scaling strong --global-batch G divides a FIXED batch of G molecules over the ranks (BASELINE
configs[4]: 8192 molecules over 2/4/8 GPUs); the default (weak) keeps --batch molecules per GPU.
At N=1 the default run also appends `workloads`: a 3-step timing of BASELINE configs[2]
(rMD17-aspirin-shape, 4096 molecules, 6 layers, Atomwise energy + forces) and configs[3]
(MD22-shape, 64 molecules of 370 atoms, lmax=3, K=160), each with its own algorithmic-byte / FLOP
fractions.
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


def algorithmic_totals(N, E, C, lmax, n_layers, sep=True):
    """(fp32 forward bytes, forward FLOPs) summed over the layers; fwd+bwd = 3x (agreed convention)."""
    L = (lmax + 1) ** 2 - 1
    S = 3 + (2 * (lmax - 1) if sep else 0)
    by = fl = 0
    for i in range(n_layers):
        last = i == n_layers - 1
        by += 4 * (2 * N * (1 + L) * C + E * C * (1 if last else 2) + E * (L + 1)) + 16 * E \
            + 4 * C * C * (10 + 3 * S + (0 if last else 2 + lmax))
        fl += N * C * C * (16 + 4 * S + 2 * L + (0 if last else 4 * L)) + E * C * C * (2 + 2 * S + (0 if last else 2))
    return by, fl


def algorithmic_bytes_per_molecule(N, E, n_mol):
    """fp32 forward bytes per layer, x3 for forward+backward."""
    by, _ = algorithmic_totals(N, E, MODEL["n_atom_basis"], MODEL["lmax"], MODEL["n_interactions"])
    return 3 * by / n_mol


def gemm_flops(M, N, K):
    return 2.0 * M * N * K



def kernel_report(prof, n_steps, N, E, peaks, ms_step):
    """Roofline of the dominant kernel + a per-entry-point table, from CUDA events recorded around every C-ABI call.
    Algorithmic bytes / FLOPs per call follow SURVEY.md §8(d) restricted to each kernel's true inputs and outputs
    (DESIGN.md §4): node arrays once, per-edge arrays once, intermediates that stay on chip count as zero."""
    C, L_, S, H = MODEL["n_atom_basis"], 8, 5, MODEL["num_heads"]
    # B200_PROFILING.md: measured peaks from MEASURED_PEAKS.json, else the stated fallbacks (6.65 TB/s; 1.59 PF burst,
    # ~1.4 PF sustained under the power cap - the step is seconds long, so the sustained figure applies)
    hbm = peaks.get("hbm_gbs", 6650.0)
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    f = 4.0
    node = N * C * f
    alg_bytes = {  # per call
        "goten_gata_fwd": E * (S + 1) * C * f + (2 + 2 * S + 2 * (1 + L_)) * node + E * (H + L_ + 3) * f,
        "goten_gata_bwd_tgt": E * C * f + E * (S + 1) * C * f + ((1 + L_) + L_ + 2 * S + 2 + 1) * node + E * (2 * H + L_ + 3) * f,
        "goten_gata_bwd_src": E * (S + 1) * C * f + ((1 + L_) + L_ + 2 * S + 1 + 2 * S + 1 + L_) * node + E * (2 * H + L_ + 3) * f,
        "goten_htr_fwd": 3 * E * C * f + 2 * L_ * node + E * L_ * f,
        "goten_htr_bwd_tgt": 3 * E * C * f + 3 * L_ * node + E * L_ * f,
        "goten_htr_bwd_src": 2 * E * C * f + 2 * L_ * node + E * L_ * f,
    }
    agg, gemm_shapes, absmax_sizes = {}, {}, {}
    for name, a, e0, e1 in prof:
        ms = e0.elapsed_time(e1)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
        if name == "goten_absmax":
            key = int(a[2]) * int(a[3])
            g = absmax_sizes.setdefault(key, [0, 0.0])
            g[0] += 1
            g[1] += ms
        if name == "goten_gemm_scaled":
            key = (int(a[8]), int(a[9]), int(a[10]), int(a[2]), int(a[5]))  # M, N, K, trans_a, trans_b
            g = gemm_shapes.setdefault(key, [0, 0.0])
            g[0] += 1
            g[1] += ms
    kernels = []
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        row = {"entry": name, "launches_per_step": cnt // n_steps, "ms_per_step": ms / n_steps,
               "share_of_step": ms / n_steps / ms_step}
        if name in alg_bytes:
            gbs = alg_bytes[name] * cnt / (ms * 1e-3) / 1e9
            row.update(bound="hbm", achieved_gbs=gbs, frac=gbs / hbm, algorithmic_mb_per_call=alg_bytes[name] / 1e6)
        elif name == "goten_gemm_scaled":
            fl = sum(2.0 * k[0] * k[1] * k[2] * c for k, (c, _) in gemm_shapes.items())
            row.update(bound="tensor", achieved_tflops=fl / (ms * 1e-3) / 1e12, frac=fl / (ms * 1e-3) / 1e12 / bf16)
        kernels.append(row)
    # dominant kernel: the tcgen05 GEMM; the roofline entry is quoted on its heaviest launch shape
    (M, Nn, K, ta, tb), (cnt, ms) = max(gemm_shapes.items(), key=lambda kv: kv[1][1])
    g_cnt, g_ms = agg["goten_gemm_scaled"]
    fl_all = sum(2.0 * k[0] * k[1] * k[2] * c for k, (c, _) in gemm_shapes.items())
    achieved = 2.0 * M * Nn * K * cnt / (ms * 1e-3) / 1e12
    arm16 = os.environ.get("GOTEN_GEMM", "auto") in ("auto", "tc16") and os.environ.get("GOTEN_TC16", "1") != "0"
    mma_per_product = 3.0 if arm16 else 6.0  # 16-bit MMA slots per fp32-accurate product
    # dram bytes per launch (read + write) from the committed ncu --set full capture of the same command (profiles/):
    # the family average when the capture lists it, else the heaviest shape's
    traffic, traffic_of = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        pre = "gemm16" if arm16 else "gemm"
        traffic, traffic_of = tj.get(f"{pre}_family_avg_per_launch"), "family average per launch"
        if traffic is None:
            traffic, traffic_of = tj.get(f"{pre}_{M}x{Nn}x{K}_{ta}{tb}"), "one launch of the heaviest shape"
    except Exception:
        pass
    family = fl_all / (g_ms * 1e-3) / 1e12
    roofline = {
        "kernel": (f"tc16::gemm16_kernel family (tcgen05 split-fp16: hi*hi + hi*lo + lo*hi, CTA pairs)" if arm16 else
                   f"tc::gemm3x_kernel family (tcgen05 3xTF32, CTA pairs)") +
                  f": all {g_cnt // n_steps} GEMM calls of a step (operand max / split passes and split-K reductions "
                  "included in the time)",
        "bound": "tensor", "achieved": family, "peak": bf16, "unit": "TFLOP/s", "frac": family / bf16,
        "traffic": traffic,
        "peak_source": "of measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else
                       "of fallback (B200_PROFILING.md: ~1.4 PF sustained bf16; HBM 6.65 TB/s)",
        "share_of_step": g_ms / n_steps / ms_step, "launches_per_step": g_cnt // n_steps,
        "flop_per_step": fl_all / n_steps,
        "heaviest_shape": {"M": M, "N": Nn, "K": K, "trans": [ta, tb], "launches_per_step": cnt // n_steps,
                           "achieved_tflops": achieved, "frac": achieved / bf16,
                           },
        "traffic_of": traffic_of,
        "note": ("fp32-accurate GEMM: operands scaled by a power of two and split x = hi + lo in fp16 (22 significant "
                 "bits), three kind::f16 MMAs per product with fp32 accumulation, so the ceiling of this kernel is "
                 "peak/3; achieved*3/peak is its tensor-pipe fraction.  FLOPs are the algorithmic 2*M*N*K of every call") if arm16 else
                ("fp32-accurate GEMM: three tf32 MMAs per product (hi*hi + hi*lo + lo*hi), each tf32 MMA costs two "
                 "bf16 MMA slots, so the ceiling of this kernel is peak/6; achieved*6/peak is its tensor-pipe fraction"),
        "tensor_pipe_frac": family * mma_per_product / bf16,
    }
    shapes = [{"M": k[0], "N": k[1], "K": k[2], "trans": [k[3], k[4]], "launches_per_step": c // n_steps,
               "ms_per_step": ms_ / n_steps, "tflops": 2.0 * k[0] * k[1] * k[2] * c / (ms_ * 1e-3) / 1e12}
              for k, (c, ms_) in sorted(gemm_shapes.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("GOTEN_BENCH_SHAPES", "14"))]]
    roofline["gemm_shapes"] = shapes
    roofline["absmax_passes"] = [{"elements": k, "launches_per_step": c // n_steps, "ms_per_step": ms_ / n_steps}
                                 for k, (c, ms_) in sorted(absmax_sizes.items(), key=lambda kv: -kv[1][1])[:10]]
    return roofline, kernels

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

    strong = args.scaling == "strong"
    if strong and args.global_batch % world:
        raise SystemExit(f"--global-batch {args.global_batch} must divide over {world} ranks")
    B = args.global_batch // world if strong else args.batch
    z, pos, batch = synth_batch("qm9", B, seed=1000 + rank)  # molecules shard by graph: each rank owns B
    zh, ph, bh = z.pin_memory(), pos.pin_memory(), batch.pin_memory()
    zd, pd, bd = z.to(dev), pos.to(dev), batch.to(dev)

    class Data:
        pass

    C_, L_out = MODEL["n_atom_basis"], (MODEL["lmax"] + 1) ** 2 - 1
    out_host = {}   # pinned host buffers for the (h, X) read-back of the inference-style e2e figure

    def step(host_inputs: bool, read_outputs: bool = False):
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
        if read_outputs:   # an (h, X) consumer: both outputs copied to pinned host memory
            if not out_host:
                out_host["h"] = torch.empty(h.shape, dtype=h.dtype).pin_memory()
                out_host["X"] = torch.empty(X.shape, dtype=X.dtype).pin_memory()
            out_host["h"].copy_(h.detach(), non_blocking=True)
            out_host["X"].copy_(X.detach(), non_blocking=True)
        if host_inputs:
            return float(loss.item())  # D2H read of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_inputs, read_outputs=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step(host_inputs, read_outputs)
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
    ms_total = timed(args.steps, False)
    launches = (L.cdll.goten_launch_count() - launches0) // max(args.steps, 1)
    for _ in range(2):
        step(True)
    ms_e2e = timed(args.steps, True)
    step(True, True)
    ms_e2e_out = timed(args.steps, True, True)
    clocks = sampler.stop() if sampler else None

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    e2e_out_value = world * B / (ms_e2e_out / args.steps * 1e-3)

    # ---- host-side enqueue time of one step (no device wait): tells how close the step is to being launch bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(False)
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    # ---- per-entry-point pass (outside the timed region): CUDA events around every C-ABI call
    n_prof = 3
    torch.cuda.synchronize()
    L.profile = []
    for _ in range(n_prof):
        step(False)
    torch.cuda.synchronize()
    prof, L.profile = L.profile, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline, kernels = kernel_report(prof, n_prof, N_nodes, E, peaks, ms_step)
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"QM9-shape synthetic, {B} molecules/GPU" +
                                   (f" (fixed global batch {args.global_batch})" if strong else "") +
                                   f" (N={N_nodes} atoms, E={E} edges on rank 0), "
                                   "cutoff 5A, max 32 nbrs, n_atom_basis=256 n_interactions=4 lmax=2 fwd+bwd "
                                   "(graph build + all parameter gradients)",
                       "parallelism": f"molecules sharded by graph over {world} GPU(s); one NCCL all-reduce of the "
                                      "flat fp32 gradient buffer" if world > 1 else "single GPU",
                       "l2": "per-step working set (~12 GB of saved activations) exceeds the 126 MB L2; no explicit flush",
                       "gemm_impl": os.environ.get("GOTEN_GEMM", "auto")},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(z.numel() * 8 + pos.numel() * 4 + batch.numel() * 8),
                    "d2h_bytes_per_step": 4,
                    "note": "training consumer: pinned-host z/pos/batch in, the scalar loss out"},
            "e2e_with_outputs": {"value": e2e_out_value, "unit": UNIT,
                                 "h2d_bytes_per_step": int(z.numel() * 8 + pos.numel() * 4 + batch.numel() * 8),
                                 "d2h_bytes_per_step": int(4 + N_nodes * C_ * 4 * (1 + L_out)),
                                 "note": "as e2e, plus h [N,C] and X [N,L,C] copied to pinned host memory every step "
                                         "(a consumer of the representation itself)"},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_ms,
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kernels,
            "algorithmic_mb_per_molecule": algorithmic_bytes_per_molecule(N_nodes, E, B) / 1e6,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(budget_s=20.0)
        if world == 1 and not args.no_workloads:
            del model, fbuf, params
            torch.cuda.empty_cache()
            out["workloads"] = {k: run_workload(k, dev, peaks) for k in ("rmd17", "md22")}
        print(json.dumps(out), file=_RESULT, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------- the other single-GPU BASELINE configurations
WORKLOADS = {
    # configs[2]: rMD17-aspirin-shape, batch 4096, 6 interactions, lmax 2, Atomwise energy head; positions require grad,
    # so pos.grad is minus the forces (first order, outputs.py:365-375).  The batch is run as 4 micro-batches of 1024
    # whose gradients accumulate (one optimiser-step's worth of work; activations of 4096 molecules would need ~120 GB)
    "rmd17": dict(kind="aspirin", batch=4096, micro=1024, max_nbr=32, head=True, model=dict(n_interactions=6, lmax=2),
                  what="BASELINE configs[2]: rMD17-aspirin-shape (21 atoms) batch=4096 as 4 accumulating micro-batches of "
                       "1024, n_atom_basis=256 n_interactions=6 lmax=2, Atomwise energy head + forces (d/dpos), fwd+bwd"),
    # configs[3]: MD22-shape, 64 molecules of 370 atoms, lmax 3, max_num_neighbors 160 (long-neighbour-list regime)
    "md22": dict(kind="md22", batch=64, micro=32, max_nbr=160, head=False, model=dict(n_interactions=4, lmax=3),
                 what="BASELINE configs[3]: MD22-shape (370 atoms, max 160 nbrs) batch=64 as 2 accumulating micro-batches "
                      "of 32, n_atom_basis=256 n_interactions=4 lmax=3, fwd+bwd"),
}


def run_workload(name, dev, peaks, steps=3):
    """3 timed steps (after 1 warm-up) of another BASELINE configuration on one GPU, with its own algorithmic-byte and
    FLOP fractions (SURVEY.md §8d formulas; fwd+bwd = 3x forward)."""
    import gotennet_b200 as g
    from gotennet_b200.synthetic import synth_batch

    w = WORKLOADS[name]
    base = {k: v for k, v in MODEL.items() if k not in ("n_interactions", "lmax")}
    torch.manual_seed(0)
    model = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(CUTOFF), max_num_neighbors=w["max_nbr"], activation="swish",
                              **base, **w["model"]).to(dev)
    head = g.Atomwise(n_in=MODEL["n_atom_basis"], n_out=1, aggregation_mode="sum", activation="swish").to(dev) \
        if w["head"] else None
    params = list(model.parameters()) + (list(head.parameters()) if head else [])
    micro = []
    for i in range(w["batch"] // w["micro"]):
        z, pos, batch = synth_batch(w["kind"], w["micro"], seed=2000 + i)
        micro.append((z.to(dev), pos.to(dev), batch.to(dev)))

    class Data:
        pass

    N_tot = E_tot = 0

    def step():
        nonlocal N_tot, E_tot
        N_tot = E_tot = 0
        for p in params:
            p.grad = None
        for zd, pd, bd in micro:
            d = Data()
            d.z, d.pos, d.batch, d.num_graphs = zd, (pd.clone().requires_grad_(True) if head is not None else pd), bd, w["micro"]
            h, X = model(d)
            if head is not None:
                d.representation, d.vector_representation = h, X
                loss = head(d)["y"].sum()
            else:
                loss = h.sum() + X.pow(2).sum()
            loss.backward()
            N_tot += model.last_plan.N
            E_tot += model.last_plan.E
        return loss

    step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    by, fl = algorithmic_totals(N_tot, E_tot, MODEL["n_atom_basis"], w["model"]["lmax"], w["model"]["n_interactions"])
    hbm = peaks.get("hbm_gbs", 6650.0)
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    res = {"workload": w["what"], "molecules": w["batch"], "atoms": N_tot, "edges": E_tot, "steps": steps,
           "ms_per_step": ms, "value": w["batch"] / (ms * 1e-3), "unit": UNIT,
           "algorithmic_gb_per_step": 3 * by / 1e9, "algorithmic_tflop_per_step": 3 * fl / 1e12,
           "hbm_frac": 3 * by / (ms * 1e-3) / 1e9 / hbm, "tensor_frac": 3 * fl / (ms * 1e-3) / 1e12 / bf16,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "finite": bool(torch.isfinite(loss).item())}
    del model, head, params, micro
    torch.cuda.empty_cache()
    return res


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


def cpu_baseline(budget_s=20.0, n_mol=64):
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
    """The reference's CPU path on this box's host cores: the oracle port (kind "port"), all host threads, on OUR arm's
    metric / config.  One step = forward + backward (graph build included) of one 64-molecule micro-batch of the
    QM9-shape workload; --warmup untimed steps, then exactly --steps timed ones."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_mol = 64
    run = _oracle_step(n_mol)
    warm, steps = max(args.warmup, 1), max(args.steps, 1)
    for _ in range(warm):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    value = n_mol / dt
    sample = (f"{n_mol}-molecule micro-batch of the QM9-shape workload per step (the 1024-molecule batch = 16 such "
              f"micro-batches), {steps} timed steps after {warm} warm-up, {dt:.2f} s per step")
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "QM9-shape synthetic, cutoff 5A, max 32 nbrs, n_atom_basis=256 n_interactions=4 lmax=2 "
                               "fwd+bwd (graph build + all parameter gradients); CPU oracle port (PyTorch restatement "
                               f"of the reference, not the verbatim package) on {cores} host threads, " + sample},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=_RESULT, flush=True)


_RESULT = sys.stdout


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1
    when NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the whole run and the result line goes to a
    private duplicate of the original stdout."""
    global _RESULT
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    _RESULT = os.fdopen(real, "w")


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="molecules per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--global-batch", type=int, default=8192, help="molecules over ALL GPUs with --scaling strong")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the rMD17 / MD22 timings appended at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
