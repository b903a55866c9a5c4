#!/usr/bin/env python
"""bench.py — scGNN forward+backward throughput (cells/sec) on synthetic cell×gene data.

One "step" = one pass of the scGNN hot path over all cells of the configuration
(BASELINE.json: "scGNN 1M cells × 2000 genes"):
  1. one Feature-AE epoch over every cell (batch 12 800, LTMG-mode loss with the reference
     driver's all-zero TRS, Adam)                                  scgnn2.py:275-335, 1217-1295
  2. one Graph-AE (GCN branch) full-batch epoch on the 128-d embeddings over the prebuilt
     k=15 kNN graph: 2 projection GEMMs + 2 SpMM forward, the exact matrix-free N×N
     inner-product-decoder loss (pos-weighted BCE + KLD), backward, Adam   scgnn2.py:575-615
kNN search + graph assembly are outside the step (the reference rebuilds the graph once per
200 Graph-AE epochs) and are reported as `graph_build_s`.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     (CPU arm: oracle port on host cores)

Prints ONE JSON line (see the repository README / DESIGN.md for the field definitions).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH = 12800
K_NN = 15
EMB = 16


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------ synthetic data
def synth_expression_gpu(n, g, device, seed=0, density=0.10, chunk=65536):
    """Device-side generator of the SURVEY §8d expression matrix (NB counts, 10 latent types,
    ~10 % non-zeros, normalize_total(1e4)+log1p through the library's own kernel)."""
    from dance_b200 import ops
    gen = torch.Generator(device=device).manual_seed(seed)
    mu_g = torch.exp(torch.randn(g, device=device, generator=gen))
    shift = torch.ones(10, g, device=device)
    for t in range(10):
        idx = torch.randperm(g, device=device, generator=gen)[:max(1, g // 20)]
        shift[t, idx] = 4.0
    X = torch.empty(n, g, dtype=torch.float32, device=device)
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        m = i1 - i0
        s_c = torch.exp(0.5 * torch.randn(m, 1, device=device, generator=gen))
        types = torch.randint(0, 10, (m, ), device=device, generator=gen)
        mean = s_c * mu_g[None, :] * shift[types]
        lam = torch._standard_gamma(torch.full_like(mean, 2.0)) * (mean / 2.0)   # NB = Poisson(Gamma), dispersion 0.5
        cnt = torch.poisson(lam, generator=gen)
        nz = (cnt > 0).float().mean().clamp_min(1e-6)
        keep = torch.rand(cnt.shape, device=device, generator=gen) < (density / nz).clamp(max=1.0)
        X[i0:i1] = cnt * keep
    ops.normalize_total_log1p_(X, target_sum=1e4, max_fraction=1.0)
    return X


# ------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    Q_OLD = Q.replace("clocks_event_reasons", "clocks_throttle_reasons")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            q = self.Q
            probe = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                   capture_output=True, text=True, timeout=20)
            if probe.returncode != 0 or "not a valid" in (probe.stdout + probe.stderr).lower():
                q = self.Q_OLD
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU (reference) arm
def cpu_reference_step_factory(n_cells, genes, seed=0):
    """The reference's CPU arithmetic for one step on `n_cells` cells: oracle.port (torch-CPU
    restatement, pinned against the reference's own code) — Feature_AE epoch + Graph_AE GCN epoch
    with the dense N×N decoder/labels exactly like scgnn2.py:557,425,603-615."""
    from oracle import port
    torch.set_num_threads(os.cpu_count())
    X = torch.from_numpy(port.synthetic_expression(n_cells, genes, seed=seed))
    torch.manual_seed(seed)
    fae = port.FeatureAE(genes)
    opt = torch.optim.Adam(fae.parameters(), lr=1e-3)
    # graph on the initial embedding (outside the timed step, like the GPU arm)
    with torch.no_grad():
        z0, _ = fae(X)
    adj, _ = port.feature2adj(z0.numpy(), K_NN)
    an = port.preprocess_graph(adj)
    adj_t = port.to_torch_sparse(an)
    pw, norm = port.gae_norm_constants(adj)
    import scipy.sparse as sp
    labels = torch.from_numpy((adj + sp.eye(n_cells)).toarray()).float()
    w = [torch.nn.Parameter(torch.empty(128, 32)), torch.nn.Parameter(torch.empty(32, EMB)), torch.nn.Parameter(torch.empty(32, EMB))]
    for p in w:
        torch.nn.init.xavier_uniform_(p)
    gopt = torch.optim.Adam(w, lr=1e-2)

    def step():
        _, z_all, _ = port.feature_ae_epoch(fae, opt, X, BATCH, "LTMG", 0.9)
        gopt.zero_grad()
        eps = torch.randn(n_cells, EMB)
        z, mu, logvar, _ = port.graph_ae_gcn_forward(z_all, w[0], w[1], w[2], adj_t, eps)
        loss = port.gae_loss(torch.mm(z, z.t()), labels, mu, logvar, n_cells, norm, pw)
        loss.backward()
        gopt.step()
        return float(loss.item())

    return step


def time_cpu(n_cells, genes, steps, warmup):
    """Times the CPU arm with the best of a few intra-op thread counts: on a 128-core host torch-CPU is ~10× SLOWER
    with 128 threads than with 32 for these shapes (measured), and the baseline should be the reference at its best."""
    step = cpu_reference_step_factory(n_cells, genes)
    cores = os.cpu_count()
    best_t, best_th = None, cores
    for th in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(th)
        step()                      # doubles as warm-up
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_th = dt, th
    torch.set_num_threads(best_th)
    time_cpu.threads = best_th
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), ts


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = args.cpu_cells
    med, ts = time_cpu(n_cpu, args.genes, args.steps, max(1, min(args.warmup, 1)))
    val = n_cpu / med
    line = {
        "impl": "reference", "metric": "cells/sec fwd+bwd scGNN", "value": val, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, loss_mode="exact-dense"),
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": time_cpu.threads, "host_cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n_cpu} of {args.cells} cells × {args.genes} genes, 1 step = Feature_AE epoch + Graph_AE GCN epoch "
                                   f"with the reference's dense {n_cpu}×{n_cpu} decoder (scgnn2.py:425,557); oracle/port.py on torch-CPU"},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, loss_mode):
    return {"workload": f"scGNN {args.cells} cells × {args.genes} genes: Feature_AE epoch (batch {BATCH}) + Graph_AE GCN epoch, k={K_NN} kNN graph",
            "cells": args.cells, "genes": args.genes, "feature_ae_batch": BATCH, "knn_k": K_NN, "graph_ae_embedding": EMB,
            "decoder_loss": loss_mode, "gemm_precision": args.precision,
            "l2_policy": "inputs larger than L2 (X is %.1f GB per rank-shard; every step streams it)" % (args.cells * args.genes * 4 / 1e9)}


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--precision", type=str, default="tf32x3", choices=["tf32x3", "tf32", "fp32"])
    ap.add_argument("--cpu-cells", type=int, default=16384, help="bounded CPU sample (cells) for cpu_baseline / --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cuda-profiler", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args)
        return

    from dance_b200 import ops
    from dance_b200.engine import FeatureAEEngine, GraphAEEngine
    from dance_b200.parallel import Comm, shard_bounds

    comm = Comm.from_env()
    rank, world = comm.rank, comm.world
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()
    ops.set_default_precision(args.precision)

    N, G = args.cells, args.genes
    bounds = shard_bounds(N, world)
    r0, r1 = bounds[rank]
    n_loc = r1 - r0

    # ---- setup (outside the timed region) ------------------------------------------------
    X = synth_expression_gpu(n_loc, G, dev, seed=1234 + rank)
    fae = FeatureAEEngine(G, device=dev, lr=1e-3, precision=args.precision, seed=0)
    gae = GraphAEEngine(128, EMB, device=dev, lr=1e-2, precision=args.precision, seed=1)
    if comm.enabled:
        fae.grad_hook = comm.allreduce_sum_
        gae.set_sharding(comm, bounds)
    z_all = torch.empty(n_loc, 128, dtype=torch.float32, device=dev)
    fae.train_epoch(X, BATCH, "LTMG", 0.9, None, z_all, None)   # also serves as the first warm-up epoch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z_full = comm.all_gather_rows(z_all, bounds) if comm.enabled else z_all
    idx_loc, _ = ops.knn(z_full, K_NN, include_rank0=False, q_begin=r0, q_end=r1, return_dist=False)
    idx_full = comm.all_gather_rows(idx_loc, bounds) if comm.enabled else idx_loc
    A_full = ops.knn_graph_build(idx_full.contiguous())
    torch.cuda.synchronize()
    graph_build_s = time.perf_counter() - t0
    nnz_total = A_full.nnz
    if comm.enabled:   # local row block of Â (column ids stay global)
        rp = A_full.rowptr[r0:r1 + 1].clone()
        e0, e1 = int(rp[0].item()), int(rp[-1].item())
        A = ops.CSR((rp - e0).contiguous(), A_full.colidx[e0:e1].contiguous(), A_full.vals[e0:e1].contiguous(), (n_loc, N))
        del A_full
    else:
        A = A_full
    labels = ops.CSR(A.rowptr, A.colidx, None, A.shape)
    adj_sum = nnz_total - N
    pos_weight = float(N * N - adj_sum) / adj_sum
    norm = N * N / float((N * N - adj_sum) * 2)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    eps = torch.empty(n_loc, EMB, dtype=torch.float32, device=dev)

    ops.reset_counters()

    def step(x_src):
        fae.train_epoch(x_src, BATCH, "LTMG", 0.9, None, z_all, None)
        eps.normal_(generator=gen)
        gae.train_step(z_all, A, labels, norm, pos_weight, eps)

    def timed(fn, k):
        comm.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        torch.cuda.synchronize()
        comm.barrier()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        comm.allreduce_max_(t)
        return float(t.item())

    for _ in range(args.warmup):
        step(X)
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs -----------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_before = ops.counters()["launches"]
    if args.cuda_profiler:
        torch.cuda.profiler.start()
    total_ms = timed(lambda: step(X), args.steps)
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    launches = ops.counters()["launches"] - launches_before
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = N / (ms_per_step / 1e3)

    # ---- per-kernel timing pass (CUDA events around every launch; same work, separate pass) ----
    ops.enable_kernel_timing(True)
    step(X)
    torch.cuda.synchronize()
    ktimes = ops.kernel_times()
    ops.enable_kernel_timing(False)

    # ---- e2e: host-resident X, pinned, H2D per batch inside the timed region --------------------
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((n_loc, G), dtype=torch.float32, pin_memory=True)
        Xh.copy_(X)
        stage = [torch.empty((BATCH, G), dtype=torch.float32, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        loss_host = torch.empty(2, dtype=torch.float32, pin_memory=True)
        nb = (n_loc + BATCH - 1) // BATCH

        def e2e_step():
            # double-buffered: batch b+1 is copied on the side stream while batch b trains
            evs = [None, None]
            free = [None, None]
            fae.loss_acc.zero_()

            def issue(b):
                buf = b & 1
                b0, b1 = b * BATCH, min(n_loc, (b + 1) * BATCH)
                with torch.cuda.stream(copy_stream):
                    if free[buf] is not None:
                        copy_stream.wait_event(free[buf])
                    stage[buf][:b1 - b0].copy_(Xh[b0:b1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    evs[buf] = ev
            issue(0)
            for b in range(nb):
                buf = b & 1
                if b + 1 < nb:
                    issue(b + 1)
                b0, b1 = b * BATCH, min(n_loc, (b + 1) * BATCH)
                main_stream.wait_event(evs[buf])
                z, _ = fae.train_step(stage[buf][:b1 - b0], None, 0.9, "LTMG")
                z_all[b0:b1].copy_(z)
                fr = torch.cuda.Event()
                fr.record(main_stream)
                free[buf] = fr
            eps.normal_(generator=gen)
            gae.train_step(z_all, A, labels, norm, pos_weight, eps)
            loss_host[0:1].copy_(fae.loss_acc, non_blocking=True)
            loss_host[1:2].copy_(gae.loss, non_blocking=True)

        e2e_step()
        e2e_ms = timed(e2e_step, args.steps) / args.steps
        e2e = {"value": N / (e2e_ms / 1e3), "unit": "cells/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(n_loc * G * 4) * world, "d2h_bytes_per_step": 8 * world,
               "note": "X pinned on the host, copied batch-by-batch on a side stream (double-buffered) inside the timed region; "
                       "per-step D2H = the two loss scalars"}
        del Xh

    if rank != 0:
        return

    # ---- roofline of the kernels (algorithmic bytes / flops from DESIGN.md §kernels) ------------
    def kt(name):
        d = ktimes.get(name)
        return (d["ms"] / d["n"], d["n"], d["ms"]) if d else (None, 0, 0.0)

    F = 32
    spmm_ms, spmm_n, spmm_total = kt("spmm_csr_f32")
    nnz_loc = A.nnz
    spmm_bytes = nnz_loc * 8 + (n_loc + 1) * 4 + N * F * 4 + n_loc * F * 4
    roof_spmm = None
    if spmm_ms:
        ach = spmm_bytes / (spmm_ms * 1e-3) / 1e9
        roof_spmm = {"kernel": "spmm_csr_kernel (Â·support, F=32)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "traffic": None, "launches": spmm_n, "ms_per_launch": spmm_ms,
                     "algorithmic_bytes": spmm_bytes, "peak_source": peaks["source"]}
    gemm_ms, gemm_n, gemm_total = kt("gemm_f32")
    nb_loc = (n_loc + BATCH - 1) // BATCH
    fae_flops = 6.0 * n_loc * (G * 512 + 512 * 128 + 128 * 512 + 512 * G)       # fwd + dX + dW, 2 flops per MAC
    gcn_flops = 6.0 * n_loc * (128 * 32 + 32 * 2 * EMB)
    roof_gemm = None
    if gemm_ms:
        ach = (fae_flops + gcn_flops) / (gemm_total * 1e-3) / 1e12
        roof_gemm = {"kernel": "gemm_tc_kernel (tcgen05 kind::tf32, %s)" % args.precision, "bound": "tensor", "achieved": ach,
                     "peak": peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": ach / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]), "traffic": None, "launches": gemm_n,
                     "ms_total": gemm_total, "algorithmic_flops": fae_flops + gcn_flops, "peak_source": peaks["source"],
                     "note": "achieved = fp32-equivalent algorithmic FLOPs; tf32x3 issues 3 tensor-core products per algorithmic product "
                             "and kind::tf32 runs at half the bf16 rate, so the ceiling of this mode is peak/6"}
    dec_ms, dec_n, dec_total = kt("gae_loss_grad_f32")
    phases = {k: {"ms": v["ms"], "launches": v["n"]} for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1]["ms"])}
    dominant = max(ktimes.items(), key=lambda kv: kv[1]["ms"])[0] if ktimes else None
    roofline = roof_gemm if dominant == "gemm_f32" else (roof_spmm if dominant == "spmm_csr_f32" else None)
    if roofline is None:
        # the exact N×N decoder is SFU/FMA-bound; report it on the tensor roofline with its matmul flops (2·d per logit, S and G·Z)
        dec_flops = 2.0 * 2 * EMB * float(n_loc) * N
        ach = dec_flops / (dec_total * 1e-3) / 1e12 if dec_total else 0.0
        roofline = {"kernel": "gae_allpairs_tch_kernel (matrix-free z·zᵀ BCE decoder: tcgen05 kind::f16 hi/lo split, S and G in TMEM)",
                    "bound": "tensor", "achieved": ach,
                    "peak": peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]), "traffic": None,
                    "launches": dec_n, "ms_total": dec_total, "peak_source": peaks["source"],
                    "note": "dominant kernel of the step; algorithmic flops = the two K=16 products per logit (S and G·Z). Its real ceiling is "
                            "the per-logit elementwise work (2 MUFU + 9 ALU instructions on 16 warps, ncu: issue 58 %, XU 58 %), not the tensor "
                            "pipe (10 % active) or HBM — see DESIGN.md §decoder and profiles/; roofline_spmm / roofline_gemm are the HBM- and "
                            "tensor-bound kernels"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        med, ts = time_cpu(args.cpu_cells, G, steps=2, warmup=1)
        cpu_baseline = {"value": args.cpu_cells / med, "unit": "cells/s", "cores": time_cpu.threads, "host_cores": os.cpu_count(), "kind": "port",
                        "sample": f"{args.cpu_cells} of {N} cells × {G} genes; same step (Feature_AE epoch + Graph_AE GCN epoch) with the "
                                  f"reference's dense {args.cpu_cells}² decoder; oracle/port.py on torch-CPU, best of 16/32/64/all intra-op threads = {time_cpu.threads}; "
                                  f"the O(N²) decoder makes per-cell CPU cost at the full {N} cells ≈{N // args.cpu_cells}× higher than in this sample",
                        "s_per_step": med}

    line = {
        "metric": "cells/sec fwd+bwd scGNN", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (tcgen05 3xTF32 split GEMMs, fp32 everything else)" if args.precision == "tf32x3" else args.precision,
        "data": "synthetic",
        "config": dict(workload_config(args, "exact-fused-blockwise (matrix-free, no N×N tensors)"), parallelism=f"cells sharded ×{world}",
                       nnz=nnz_total),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches // args.steps,
        "roofline": roofline, "roofline_spmm": roof_spmm, "roofline_gemm": roof_gemm, "cpu_baseline": cpu_baseline,
        "kernel_ms_per_step": phases, "graph_build_s": graph_build_s,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
