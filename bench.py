#!/usr/bin/env python
"""bench.py -- pairwise OT distances / second on the PILOT patient-distance hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

One "step" = one pass of the whole hot path (proportion histogram -> median centroids -> cdist
-> all-pairs OT -> dense S x S matrix) over one synthetic cohort.  The default workload is
BASELINE.json configs[1] ("c2": 1M cells, 50-dim embedding, 30 cell types, 100 samples,
stabilised Sinkhorn reg = 0.1 on one B200).  Prints ONE JSON line (rank 0).

  value     -- problems/s with the inputs already resident in HBM (CUDA-event timed)
  e2e       -- the same metric through the public API pilot_b200.tl.wasserstein_distance(adata)
               with HOST inputs (pinned embedding), H2D/D2H inside the timed region
  roofline  -- the dominant kernel of the step against the measured peak
  cpu_baseline -- the oracle port of the reference's CPU path on the box's host cores (N = 1 only)
  kernels   -- extra: C5-shaped (K = 64, 20 000 samples) slices of the two pair kernels

N > 1 (torchrun, one rank per GPU): weak scaling -- the cohort grows so every rank keeps 10^4 OT
problems (S_N = ceil(100 sqrt(N))), the pair space is partitioned over the ranks and assembled
with one NCCL all-gather inside the timed region; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pairwise OT distances/sec (exact EMD & Sinkhorn) at 1/2/4/8 B200"
UNIT = "OT problems/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def workload_shape(name: str, n_gpus: int):
    from pilot_b200 import synth
    n, d, k, s, seed = synth.CONFIGS[name]
    s_n = int(math.ceil(s * math.sqrt(n_gpus))) if n_gpus > 1 else s
    return n, d, k, s_n, seed


def make_workload(name: str, n_gpus: int):
    from pilot_b200 import synth
    n, d, k, s, seed = workload_shape(name, n_gpus)
    X, obs = synth.make_cells(n, d, k, s, seed, labels="categorical")
    return X, obs, (n, d, k, s)


REG = {"c1": None, "c2": 0.1, "c3": 0.1, "c4": 0.01}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(X, obs, reg, rows):
    """One bounded pass of the reference's CPU path (oracle port; the reference itself is verbatim-
    exec'd when /root/reference is mounted): stages 1-2 in full, stage 3 on the first `rows` rows
    of the ordered pair matrix through the reference's Python loop semantics (NumPy
    sinkhorn_stabilized restatement / C network simplex, one core, per-call overhead included)."""
    import pandas as pd
    from oracle import pilot_oracle as po
    from oracle import ref_exec
    annot = obs[["cell_types", "sampleID", "status"]].copy()
    annot.columns = ["cell_type", "sampleID", "status"]
    data = pd.DataFrame(X)
    t0 = time.perf_counter()
    if ref_exec.available():
        ref = ref_exec.load()
        props = ref.Cluster_Representations(annot)
        t1 = time.perf_counter()
        dis, _ = ref.cost_matrix(annot, data, "cosine")
    else:
        props = po.cluster_representations(annot)
        t1 = time.perf_counter()
        dis, _ = po.cost_matrix(annot, data, "cosine")
    t2 = time.perf_counter()
    ids = list(props.keys())
    S = len(ids)
    M = dis / dis.max()
    rows = min(rows, S)
    out = np.zeros((rows, S))
    for i in range(rows):
        for j in range(S):
            if reg is None:
                out[i, j] = po.emd2(props[ids[i]], props[ids[j]], M)
            else:
                out[i, j] = po.sinkhorn2_np(props[ids[i]], props[ids[j]], M, reg)
    t3 = time.perf_counter()
    full = (t1 - t0) + (t2 - t1) + (t3 - t2) * S / rows
    return dict(t_props=t1 - t0, t_cost=t2 - t1, t_pairs_sample=t3 - t2, rows=rows, S=S,
                t_step_extrapolated=full, problems_per_s=S * S / full,
                kind="reference+port" if ref_exec.available() else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, obs, (n, d, k, s) = make_workload(args.workload, 1)
    reg = REG[args.workload]
    rows = 4
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            for _ in range(args.warmup):
                cpu_reference_step(X, obs, reg, 1)
            t0 = time.perf_counter()
            res = [cpu_reference_step(X, obs, reg, rows) for _ in range(args.steps)]
            wall = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    step = float(np.mean([r["t_step_extrapolated"] for r in res]))
    value = s * s / step
    sample = (f"stages 1-2 in full ({n} cells); stage 3 on the first {rows} of {s} rows "
              f"({rows * s} ordered problems, NumPy sinkhorn_stabilized restatement in the reference's Python "
              f"loop), extrapolated x{s}/{rows}; measured wall per step {wall / args.steps:.2f}s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n} cells x {d} dims, {k} types, {s} samples, "
                                   f"Sinkhorn reg={reg}" if reg else f"{args.workload}: exact EMD"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": res[0]["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class DeviceStep:
    """The hot path with device-resident inputs, stage by stage (what tl.wasserstein_distance runs)."""

    def __init__(self, X, obs, reg):
        import torch
        from pilot_b200 import tl
        self.torch = torch
        self.reg = reg
        annot = obs[["cell_types", "sampleID", "status"]].copy()
        annot.columns = ["cell_type", "sampleID", "status"]
        self.lab = tl._Labels(annot, "cell_type", "sampleID")      # codes on device + perms
        self.X = tl._embedding_to_device(X)
        self.launches = 0

    def run(self, timers=None):
        from pilot_b200 import ops, pairs
        torch = self.torch
        lab = self.lab

        def mark(name):
            if timers is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                timers.append((name, ev))

        mark("start")
        counts_raw, first_ct, first_smp = ops.hist(lab.ct_dev, lab.sm_dev, lab.K_raw, lab.S_raw)
        mark("hist")
        props, counts = ops.props_finalize(counts_raw, lab.perm_k_dev, lab.perm_s_dev, lab.n, 0.2, True)
        mark("props")
        _, cent64_raw = ops.centroid_median(self.X, lab.ct_dev, lab.K_raw)
        mark("median")
        cent64 = cent64_raw.index_select(0, lab.perm_k_dev.long()).contiguous()
        cost, cost_norm, _ = ops.cdist(cent64, "cosine")
        mark("cdist")
        dense = pairs.all_pairs(props, cost_norm, "unreg" if self.reg is None else "reg",
                                self.reg if self.reg is not None else 0.1)
        mark("pairs")
        return dense, props, cost


# kernel launches of ONE DeviceStep.run() (my kernels only; memsets and torch's index_select excluded):
# hist_init + hist (2), prior + finalize (2), median count + plan + scatter + pivot + stream + finish (6),
# cdist prep/pair/norm (3), unpack (1) and the pair stage: emd (1), or Sinkhorn setup + reference-form redo (2) +
# for K <= 32 the warp solver + the general panel and tail variants that return at once when the cost is
# symmetric (3), else both variants of the panel and of the tail kernel (4)
def launches_per_step(reg, k=30):
    pair = 1 if reg is None else (2 + (3 if k <= 32 else 4))
    return 2 + 2 + 6 + 3 + pair + 1


def pair_kernel_slices(peak_fp64):
    """C5-shaped slices (K = 64, S = 20 000): the two pair kernels alone, device resident."""
    import torch
    from oracle import pilot_oracle as po
    from pilot_b200 import _lib, ops, synth
    S, K = 20_000, 64
    P, M = synth.make_pairs(S, K, seed=5)
    Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()
    res = {}

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best, out

    # Sinkhorn: first 96 rows x 20 000 columns (ordered problems)
    rows = 96
    rng = ops.make_range(rows * S, _lib.PAIRS_FULL)
    ms, out = timed(lambda: ops.sinkhorn_pairs(Pd, Md, 0.1, rng, want_info=True))
    iters = out[1].sum().item()
    flops = float(iters) * 4 * K * K
    res["sinkhorn_c5_slice"] = {
        "problems": rows * S, "ms": ms, "problems_per_s": rows * S / (ms * 1e-3),
        "mean_iters": iters / (rows * S), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
        "fp64_fma_peak_tflops": peak_fp64, "frac_of_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak_fp64}
    # exact EMD: first 1 000 000 pairs of the upper triangle
    n_emd = 1_000_000
    rng = ops.make_range(n_emd, _lib.PAIRS_UPPER)
    rng.total = n_emd
    ms, out = timed(lambda: ops.emd_pairs(Pd, Md, rng, want_info=True))
    # algorithmic work of the REFERENCE algorithm on the same kind of input: 3A + 2U + 2C
    _, _, st = po.emd_rows(P[:24], M, 0, 24, return_stats=True)
    ops_per_problem = (3 * st["arcs_priced"] + 2 * st["pot_updates"] + 2 * st["cycle_steps"]) / (24 * 24)
    res["emd_c5_slice"] = {
        "problems": n_emd, "ms": ms, "pairs_per_s": n_emd / (ms * 1e-3),
        "mean_pivots": out[2].float().mean().item(),
        "reference_algorithm_fp64_ops_per_problem": ops_per_problem,
        "algorithmic_tflops": ops_per_problem * n_emd / (ms * 1e-3) / 1e12,
        "fp64_fma_peak_tflops": peak_fp64,
        "frac_of_fp64_peak": ops_per_problem * n_emd / (ms * 1e-3) / 1e12 / peak_fp64}
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pilot_b200 import ops, synth, tl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator comes up; the contract is
        # ONE JSON line on stdout, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n_gpus = world

    X, obs, (n, d, k, s) = make_workload(args.workload, n_gpus)
    reg = REG[args.workload]
    hbm_peak, peak_kind = load_peaks()

    tmp = tempfile.mkdtemp()
    os.chdir(tmp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ("value") ----------------
    step = DeviceStep(X, obs, reg)
    for _ in range(args.warmup):
        step.run()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dense, props, cost = step.run()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = s * s / (ms_step * 1e-3)

    # ---------------- per-kernel breakdown + roofline of the dominant kernel ----------------
    stage_ms = {}
    for _ in range(3):
        timers = []
        step.run(timers)
        torch.cuda.synchronize()
        for (n0, ev0), (n1, ev1) in zip(timers[:-1], timers[1:]):
            stage_ms.setdefault(n1, []).append(ev0.elapsed_time(ev1))
    stage_ms = {kname: float(np.mean(v)) for kname, v in stage_ms.items()}
    elt = X.dtype.itemsize
    alg_bytes = {"hist": n * 8 + s * k * 8, "median": n * d * elt + n * 4 + k * d * elt}
    dominant = max(stage_ms, key=stage_ms.get)
    # FP64 pipe peaks are not in MEASURED_PEAKS.json: measure them here (DFMA, FFMA, FP64 mma.sync)
    peaks = {"fp64_fma_tflops": ops.pipe_peak(0), "fp32_fma_tflops": ops.pipe_peak(1),
             "fp64_dmma_tflops": ops.pipe_peak(2)}
    if dominant in alg_bytes:
        ach = alg_bytes[dominant] / (stage_ms[dominant] * 1e-3) / 1e9
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                    "algorithmic_bytes": alg_bytes[dominant], "ms": stage_ms[dominant]}
    else:
        # the all-pairs stage dominates: algorithmic flops = sum over this rank's problems of iters * 4 K^2
        # (SURVEY 8d); with K <= 32 the matvecs are DFMA chains of sinkhorn_warp_kernel (one warp per problem),
        # above that DMMA panels of sinkhorn_batched_kernel -- same FP64 pipe, same peak
        from pilot_b200 import _lib, pairs as _pairs
        if reg is not None:
            total = s * s
            rng = _lib.PairRange(total=total, block=_pairs.choose_block(total, world), nranks=world, rank=rank,
                                 mode=_lib.PAIRS_FULL, reserved=0)
            _, it, _, _ = ops.sinkhorn_pairs(props, cost / cost.max(), reg, rng, want_info=True)
            flops = float(it.sum().item()) * 4.0 * k * k
            mean_iters = float(it.float().mean().item())
            max_iters = int(it.max().item())
        else:
            flops, mean_iters, max_iters = float("nan"), None, None
        ach = flops / (stage_ms[dominant] * 1e-3) / 1e12
        kern = "sinkhorn_warp_kernel" if k <= 32 else "sinkhorn_batched_kernel"
        roofline = {"kernel": kern + " (all-pairs stage incl. setup/unpack)", "bound": "tensor",
                    "achieved": ach, "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["fp64_dmma_tflops"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at the
                    # C2 pair stage (profiles/sinkhorn_warp_r1.txt): the inputs are 24 KB of proportions and a 7 KB cost
                    "traffic": 62720 if (k <= 32 and args.workload == "c2") else None,
                    "peak_source": "FP64 mma.sync peak measured in this run by pilot_pipe_peak (MEASURED_PEAKS.json "
                                   "holds only bf16 and HBM peaks)",
                    "algorithmic_flops": flops, "mean_iters": mean_iters, "max_iters": max_iters,
                    "ms": stage_ms[dominant],
                    "note": "C2 is 10^4 problems of up to 1000 dependent iterations: latency-bound by construction; "
                            "see kernels.sinkhorn_c5_slice for the throughput-bound figure"}

    # ---------------- end to end through the public API ----------------
    pinned = torch.empty(X.shape, dtype=torch.float32 if X.dtype == np.float32 else torch.float64, pin_memory=True)
    Xp = pinned.numpy()
    Xp[...] = X
    kw = dict(emb_matrix="X_PCA", clusters_col="cell_types", sample_col="sampleID", status="status",
              regularized="unreg" if reg is None else "reg", reg=reg if reg is not None else 0.1)
    def api_step():
        adata = synth.FakeAnnData(obs, obsm={"X_PCA": Xp})
        tl.wasserstein_distance(adata, **kw)
        return adata
    for _ in range(max(1, min(args.warmup, 3))):
        api_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        adata = api_step()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = te.item()
    # embedding + the two label-code columns (categorical obs: int8 codes for < 128 categories, else int16/32)
    code_bytes = sum(1 if c < 128 else (2 if c < 32768 else 4) for c in (k, s))
    h2d = n * code_bytes + X.nbytes + (k + s) * 4
    d2h = s * s * 8 + s * k * 8 + k * k * 8 + (k + s) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} (BASELINE configs[1]): {n} cells x {d}-dim {X.dtype} embedding, "
                                   f"{k} cell types, {s} samples, cosine cost, "
                                   + (f"stabilised Sinkhorn reg={reg}, all {s * s} ordered pairs" if reg is not None
                                      else "exact EMD"),
                       "l2": "inputs (embedding %.0f MB) larger than the 126 MB L2; no flush" % (X.nbytes / 1e6),
                       "multi_gpu": "pair space block-partitioned over ranks, one NCCL all-gather per step; "
                                    "stages 1-2 replicated; samples scale as ceil(100*sqrt(N))"},
            "clocks": clocks,
            "e2e": {"value": s * s / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e * 1e3,
                    "api": "pilot_b200.tl.wasserstein_distance(adata) with categorical obs and a pinned host embedding"},
            "gpu_launches": launches_per_step(reg, k) * args.steps,
            "stage_ms": stage_ms, "roofline": roofline}

    line["pipe_peaks"] = peaks
    if n_gpus == 1:
        # the two pair kernels alone at the C5 shape (K = 64, S = 20 000)
        try:
            line["kernels"] = pair_kernel_slices(peaks["fp64_fma_tflops"])
        except Exception as exc:  # keep the headline line even if the extra slices fail
            line["kernels"] = {"error": repr(exc)}
        cb = cpu_reference_step(X, obs, reg, 4)
        line["cpu_baseline"] = {
            "value": cb["problems_per_s"], "unit": UNIT, "cores": 1, "kind": cb["kind"],
            "sample": f"stages 1-2 in full ({cb['t_props']:.2f}s + {cb['t_cost']:.2f}s); stage 3 on the first "
                      f"{cb['rows']} of {cb['S']} rows ({cb['t_pairs_sample']:.2f}s), extrapolated x{cb['S']}/{cb['rows']}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE configs[4], the scaling sweep: 20 000 synthetic samples x 64 cell types, ALL pairs with both
    solvers (exact EMD: 2.0e8 unordered pairs; Sinkhorn reg 0.1: 4.0e8 ordered problems), pair space
    partitioned over the ranks, one NCCL all-gather per matrix, dense S x S result on every rank.
    Strong scaling (the work is fixed).  A random sample of entries is checked against the CPU oracle."""
    import torch
    import torch.distributed as dist
    from pilot_b200 import _lib, ops, pairs, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    S, K, reg = 20_000, 64, 0.1
    P, M = synth.make_pairs(S, K, seed=5)
    Pd, Md = torch.from_numpy(P).cuda(), torch.from_numpy(M).cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on a slice (kernels, NCCL buffers, allocator), then the timed full matrices
    for _ in range(max(1, args.warmup)):
        pairs.all_pairs(Pd[:512], Md, "unreg")
        pairs.all_pairs(Pd[:512], Md, "reg", reg)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    times = {"emd": [], "sinkhorn": []}
    emd = sk = None
    for _ in range(max(1, args.steps)):
        for name, regularized in (("emd", "unreg"), ("sinkhorn", "reg")):
            if name == "emd":
                emd = None
            else:
                sk = None
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = pairs.all_pairs(Pd, Md, regularized, reg)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[name].append(t.item())
            if name == "emd":
                emd = res
            else:
                sk = res
    clocks = sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    from oracle import pilot_oracle as po
    rs = np.random.default_rng(0)
    ii, jj = rs.integers(0, S, 200), rs.integers(0, S, 200)
    got_e = emd[torch.from_numpy(ii).cuda(), torch.from_numpy(jj).cuda()].cpu().numpy()
    got_s = sk[torch.from_numpy(ii).cuda(), torch.from_numpy(jj).cuda()].cpu().numpy()
    want_e = np.array([po.emd2(P[i], P[j], M) for i, j in zip(ii, jj)])
    want_s = np.array([po.sinkhorn2(P[i], P[j], M, reg) for i, j in zip(ii, jj)])
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    sym = float((emd[:2000, :2000] - emd[:2000, :2000].T).abs().max().item())
    checks = {"sampled_entries": 200, "emd_max_rel_err_vs_oracle": rel(got_e, want_e),
              "sinkhorn_max_rel_err_vs_oracle": rel(got_s, want_s),
              "emd_le_sinkhorn": bool((got_e <= got_s * (1 + 1e-9)).all()),
              "emd_symmetry_abs_2000x2000": sym, "emd_diag_abs_max": float(emd.diagonal().abs().max().item())}
    n_emd, n_sk = S * (S - 1) // 2, S * S
    ms_e, ms_s = float(np.mean(times["emd"])), float(np.mean(times["sinkhorn"]))
    ms = ms_e + ms_s
    line = {"metric": METRIC, "value": (n_emd + n_sk) / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": max(1, args.steps), "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c5 (BASELINE configs[4]): 20000 samples x 64 cell types, all pairs: exact EMD "
                                   "(199990000 unordered) + stabilised Sinkhorn reg=0.1 (400000000 ordered), dense "
                                   "S x S f64 result on every rank",
                       "warmup": "512-sample slice of both solvers"},
            "clocks": clocks,
            "emd": {"pairs": n_emd, "ms": ms_e, "pairs_per_s": n_emd / (ms_e * 1e-3)},
            "sinkhorn": {"problems": n_sk, "ms": ms_s, "problems_per_s": n_sk / (ms_s * 1e-3)},
            "checks": checks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
