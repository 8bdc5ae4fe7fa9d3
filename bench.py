#!/usr/bin/env python
"""Benchmark of the spatpcaCV hot path (BASELINE.json: spatpcaCV wall time and ADMM iterations/s at
p = 10 000, n = 200, K = 5, 1-8 B200 vs host CPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

One "step" is one complete spatpcaCV job (Omega, the 57 SPD systems, every ADMM call, all CV losses,
selection of tau1 / tau2 / gamma, final fit) on the synthetic field of SURVEY.md section 8d.
`value` is ADMM iterations per second of the whole job with the inputs already resident in HBM;
`e2e` is the same through the reference-facing C-ABI call `spb_cv` with host buffers (H2D and D2H
copies inside the timed region).  Under torchrun (N > 1) the folds, the stage-1 systems and Omega's
columns of the same job are spread over the ranks ("strong" scaling) and timing is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spatpcaCV ADMM iterations/s (whole CV job: Omega + SPD inverses + ADMM + CV losses)"
UNIT = "iters/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
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
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_workload(name):
    from spatpca_b200.synth import make_case
    c = make_case(name)
    cfg = {
        "workload": f"{name}: {c['x'].shape[1]}-D {'irregular' if name in ('C3', 'C4', 'C5') else 'grid'} locations "
                    f"p={c['x'].shape[0]}, n={c['Y'].shape[0]}, K={c['K']}, "
                    f"M={c['M']}-fold CV, tau1 x tau2 x gamma = {len(c['tau1'])} x {len(c['tau2'])} x {len(c['gamma'])}, "
                    "maxit=100, thr=1e-4 (SURVEY.md section 8d generator, seed 2024)",
        "p": int(c["x"].shape[0]), "n": int(c["Y"].shape[0]), "K": int(c["K"]), "M": int(c["M"]),
        "n_tau1": len(c["tau1"]), "n_tau2": len(c["tau2"]), "n_gamma": len(c["gamma"]),
        "l2_policy": "inputs (800 MB p x p operands) exceed the 126 MB L2; no reuse across steps",
    }
    return c, cfg


def parallelism_note(M, world):
    """Part of `config` in BOTH arms (the reference arm runs the same job on the host cores of rank 0)."""
    if world <= 1:
        return "one GPU: folds run concurrently on per-fold streams"
    from spatpca_b200 import distributed as D
    return (f"folds sharded over {world} rank(s), full-data fit on rank {D.full_fit_owner(M, world)}; Omega built in "
            "column blocks (one per rank) exchanged by NCCL broadcast; CV scores all-gathered per stage")


def run_job_resident(sp, case, coll=None):
    """One spatpcaCV job with the inputs already in HBM: returns (seconds, result, stats)."""
    import torch
    from spatpca_b200 import distributed as D
    eng = case["_engine"]
    M = case["M"]
    torch.cuda.synchronize()
    if coll is not None and coll.world > 1:
        coll.barrier()
    t0 = time.perf_counter()
    if coll is None:
        coll = D.Collective()
    D.prepare_sharded(eng, coll, D.my_folds(M, coll))   # Omega shared by the ranks (column blocks over NCCL), SVD starts
    res = D.run_sharded(eng, sp.select_index, coll, M, len(case["tau1"]), len(case["tau2"]), len(case["gamma"]),
                        case["tau1"], case["tau2"], case["gamma"],
                        share_inverses=os.environ.get("SPB_SHARE_INVERSES", "1") != "0")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, res


def new_engine(sp, case):
    e = sp.Engine(case["x"].shape[0], case["x"].shape[1], case["Y"].shape[0], case["M"], case["K"],
                  case["tau1"], case["tau2"], case["gamma"], 100, 1e-4, case["l2"])
    e.upload(case["x"], case["Y"], case["nk"])
    return e


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's path) -- bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(case, omega, tau1_idx=(0, 5), fold=0):
    """Stage-1 work of one fold restricted to two tau1 values: SVD start, Gram matrix, one
    inv_sympd + its ADMM iterations (spatpcaCore2), two hold-out losses.  Returns (seconds, iters)."""
    from oracle import spatpca_ref as ref
    Y, nk, K = case["Y"], case["nk"], case["K"]
    tr = ref.Trace()
    t0 = time.perf_counter()
    Ytr, Yva = Y[nk != (fold + 1)], Y[nk == (fold + 1)]
    _, rho, Phi0, L20 = ref.pca_init(Ytr, K)
    G = np.asfortranarray(Ytr.T @ Ytr)
    loss0 = ref.cv_loss(Yva, Phi0)
    Phi, C, L2 = Phi0, Phi0, L20
    losses = [loss0]
    for i in tau1_idx[1:]:
        Phi, C, L2 = ref.core2(G, Phi, C, L2, omega, case["tau1"][i], rho, 100, 1e-4, tr)
        losses.append(ref.cv_loss(Yva, Phi))
    dt = time.perf_counter() - t0
    return dt, tr.admm_iters, tr


def reference_inverse_count(case, index):
    """inv_sympd call sites the reference executes for this job (src/RcppSpatPCA.cpp:165, 203, 397, 661, 673)."""
    M, n1, n2 = case["M"], len(case["tau1"]), len(case["tau2"])
    index1 = index[0]
    sel1 = case["tau1"][index1] if n1 > 1 else float(np.max(case["tau1"]))
    cnt = 0
    if n1 > 1:
        cnt += (n1 - 1) * M + (1 if index1 > 0 else 0)            # stage-1 paths + Core2p
    elif sel1 != 0:
        cnt += 1 + 2 * M
    if n2 > 1:
        cnt += M * ((1 if sel1 != 0 else 0) + 1) + 1              # stage-2 Core2 + tempinv per fold, full tempinv
    elif float(np.max(case["tau2"])) > 0:
        cnt += 1
    if n2 == 1 and (sel1 != 0 or float(np.max(case["tau2"])) != 0):
        cnt += M                                                   # stage-3 Core2 refits
    return cnt


def pin_blas_threads():
    """All host cores for the oracle's BLAS/LAPACK, whatever the launcher exported (torchrun forces
    OMP_NUM_THREADS=1, which made the round-1 reference arm run on one core).  Returns the thread count in use."""
    try:
        want = len(os.sched_getaffinity(0))
    except Exception:
        want = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=want, user_api="blas")
        nums = [d.get("num_threads", 1) for d in threadpoolctl.threadpool_info() if d.get("user_api") == "blas"]
        if nums:
            return int(max(nums))
    except Exception:
        pass
    return want


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        nums = [d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"]
        if nums:
            return int(max(nums))
    except Exception:
        pass
    return os.cpu_count() or 1


def oracle_full_job(case):
    """The reference's whole spatpcaCV job (oracle port) on the host: returns (seconds, result, trace)."""
    from oracle import spatpca_ref as ref
    tr = ref.Trace()
    t0 = time.perf_counter()
    res = ref.spatpca_cv(case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"],
                         case["nk"], 100, 1e-4, case["l2"], trace=tr)
    return time.perf_counter() - t0, res, tr


def reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU path on the host cores of rank 0 -- the oracle port, because the
    reference itself needs R + RcppArmadillo (absent in this image).  It runs the SAME configuration as the GPU arm:
    complete spatpcaCV jobs (Omega, all 62 inv_sympd, every ADMM call, all losses, the selections).  One job takes
    minutes at C4, so the arm runs jobs until `--steps` are done or REF_BUDGET_S (default 240 s) is spent -- at least
    one complete job -- and reports the steps it really timed."""
    if rank != 0:
        return 0
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 1))
    cores = pin_blas_threads()
    case, cfg = make_workload(args.config)
    cfg["parallelism"] = parallelism_note(case["M"], world)
    budget = float(os.environ.get("REF_BUDGET_S", "240"))
    t_start = time.perf_counter()
    times, iters, res, tr = [], 0, None, None
    while len(times) < max(1, args.steps):
        dt, res, tr = oracle_full_job(case)
        times.append(dt)
        iters = tr.admm_iters
        if time.perf_counter() - t_start + dt > budget:
            break
    mean_t = float(np.mean(times))
    value = iters / mean_t
    sample = (f"{len(times)} complete spatpcaCV job(s) of this configuration on {cores} BLAS threads "
              f"({mean_t:.1f} s per job, {iters} ADMM iterations in {len(tr.calls)} calls; no warm-up: every job is cold); "
              f"phases of the last job (s): " + ", ".join(f"{k}={v:.1f}" for k, v in sorted(tr.phase_s.items())))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": mean_t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "wall_s": mean_t, "admm_iters": iters, "admm_calls": len(tr.calls),
        "selected": [res["selected_tau1"], res["selected_tau2"], res["selected_gamma"]],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def protect_stdout():
    """Rank 0 must print ONE JSON line: libraries (NCCL prints its version banner to stdout) get stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-forced", action="store_true", help="skip the forced-iteration leg (thr = 0)")
    ap.add_argument("--forced-maxit", type=int, default=100)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import spatpca_b200 as sp
    from spatpca_b200 import _lib, engine as eng_mod
    from spatpca_b200 import distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (spatpca_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.load().spb_set_device(local_rank))
    coll = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        coll = D.TorchCollective(torch.device("cuda", local_rank))

    case, cfg = make_workload(args.config)
    cfg["parallelism"] = parallelism_note(case["M"], world)
    p, K = cfg["p"], cfg["K"]
    hbm_peak, peak_src = load_peaks()

    # ---- device-resident runs: `value` ----
    case["_engine"] = new_engine(sp, case)
    for _ in range(args.warmup):
        case["_engine"].close()
        case["_engine"] = new_engine(sp, case)
        run_job_resident(sp, case, coll)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times, stats, res = [], None, None
    for _ in range(args.steps):
        case["_engine"].close()
        case["_engine"] = new_engine(sp, case)          # upload outside the timed region
        dt, res = run_job_resident(sp, case, coll)
        times.append(dt)
        stats = case["_engine"].stats()
    clocks = sampler.stop() if rank == 0 else None
    t_local = float(np.mean(times))
    iters_local = int(stats["admm_iters"])
    launches_local = int(stats["gpu_launches"])
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([t_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_job = float(t.item())
        c = torch.tensor([iters_local, launches_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        iters_job, launches_job = int(c[0].item()), int(c[1].item())
    else:
        t_job, iters_job, launches_job = t_local, iters_local, launches_local
    value = iters_job / t_job

    # the CPU-baseline sample takes Omega from the GPU run; fetch it before the engine is released
    om_for_cpu = None
    if not args.no_cpu_baseline and world == 1:
        om_for_cpu = case["_engine"].get_omega()

    # ---- end to end through the C ABI with host buffers: `e2e` (single-GPU entry point) ----
    e2e = None
    if world == 1:
        # release the resident engine first: while it holds its ~50 GB of cached systems the e2e engine cannot
        # reuse them from the allocator cache and pays ~150 ms of cudaMalloc per job that a user never sees
        case["_engine"].close()
        args_cv = (case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"], case["nk"],
                   100, 1e-4, case["l2"])
        sp.spatpcaCV(*args_cv)                                      # warm-up
        t_e2e, st_e2e = [], None
        for _ in range(args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r2 = sp.spatpcaCV(*args_cv, return_stats=True)
            t_e2e.append(time.perf_counter() - t0)
            st_e2e = r2["stats"]
        e2e = {"value": st_e2e["admm_iters"] / float(np.mean(t_e2e)), "unit": UNIT,
               "h2d_bytes_per_step": int(st_e2e["bytes_h2d"]), "d2h_bytes_per_step": int(st_e2e["bytes_d2h"]),
               "ms_per_step": float(np.mean(t_e2e)) * 1e3,
               "selected": [r2["selected_tau1"], r2["selected_tau2"], r2["selected_gamma"]]}
    else:
        # N > 1: the public multi-GPU call is the sharded runner itself; its timed region already
        # contains the NCCL gathers and the D2H reads of every score row; add the H2D upload.
        t_up = []
        for _ in range(args.steps):
            case["_engine"].close()
            if coll is not None:
                coll.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            case["_engine"] = new_engine(sp, case)
            dt, _ = run_job_resident(sp, case, coll)
            torch.cuda.synchronize()
            t_up.append(time.perf_counter() - t0)
        import torch.distributed as dist
        t = torch.tensor([float(np.mean(t_up))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = case["_engine"].stats()
        e2e = {"value": iters_job / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(st["bytes_h2d"]) * world, "d2h_bytes_per_step": int(st["bytes_d2h"]),
               "ms_per_step": float(t.item()) * 1e3}

    if rank != 0:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- rooflines (all timed in THIS run with CUDA events on the library's stream; peaks: MEASURED_PEAKS.json) ----
    # (1) headline = the own kernel with the largest share of the timed step: batched_core2_kernel, the fold-batched
    #     matrix-free spatpcaCore2 (csrc/batched_cg.cu).  All M folds of a stage advance in lock-step; every
    #     conjugate-gradient step streams Omega_u ONCE for the M*K = 25 columns on the TMA-fed FP64 DMMA stream
    #     (cp.async.bulk.tensor + mbarrier -> mma.sync m8n8k4), the rank-n part rides the same pipeline.  With 25
    #     columns the product is FP64-tensor-bound (2 p^2 * 25 flop over 8 p^2 bytes = 6.25 flop/B), so the bound is
    #     "tensor" and the peak is the FP64 GEMM rate measured in this run (cuBLAS DGEMM 8192^3: MEASURED_PEAKS.json
    #     carries no FP64 figure).  Algorithmic flop per matrix pass: 2 p^2 M K (Omega_u) + 4 n p M K (the two p x n
    #     contractions); bytes per pass for the HBM view: 8 p^2 + 16 n p.
    n_all, Mf = cfg["n"], cfg["M"]
    n_tr = int(round(n_all * (Mf - 1) / Mf))
    roof_admm_iters = 3
    dgemm_tf = eng_mod.bench_dgemm(8192, 2)
    batched_ok = K <= 8 and Mf * K <= 32 and eng_mod.core2_batched_supported(p, n_all, K, Mf)
    stream = {}
    for label, ncols in (("single_problem", K), ("all_folds", min(32, Mf * K))):
        try:
            ms = eng_mod.bench_dmma_matvec(p, ncols, 5)
        except Exception:
            continue
        stream[label] = {"kernel": "dmma_matvec_kernel (the TMA + DMMA stream alone, U = Omega_u V)", "columns": ncols,
                         "us_per_pass": ms * 1e3, "hbm_gbs": 8.0 * p * p / (ms * 1e-3) / 1e9,
                         "hbm_frac": 8.0 * p * p / (ms * 1e-3) / 1e9 / hbm_peak,
                         "fp64_tflops": 2.0 * p * p * ncols / (ms * 1e-3) / 1e12,
                         "fp64_frac_of_dgemm": 2.0 * p * p * ncols / (ms * 1e-3) / 1e12 / dgemm_tf}
    cg_ms, cg_passes = (None, 0)
    if batched_ok:
        cg_ms, cg_passes = eng_mod.bench_core2_batched(p, n_all, K, Mf, roof_admm_iters, 3)
    cg_flop_pass = 2.0 * p * p * Mf * K + 4.0 * n_all * p * Mf * K
    cg_bytes_pass = 8.0 * p * p + 16.0 * n_all * p + 96.0 * p * K * Mf
    single_ms, single_passes = eng_mod.bench_core2_cg(p, K, n_tr, 4, 3) if K <= 8 else (None, 0)
    # (2) the persistent ADMM kernel on explicit / factored systems (the tau2-path systems of stages 2-3 and every
    #     system the matrix-free kernel is not chosen for): 8 p^2 + 72 p K bytes per Core3 iteration
    roof_iters = 10
    alg_bytes_iter = 8.0 * p * p + 72.0 * p * K                     # SURVEY.md section 8d, U-iter (Core3)
    m_path, m_stage1 = eng_mod.default_path_blocks(p), eng_mod.default_blocks(p)
    modes = {}
    for label, m in (("path_systems", m_path), ("stage1_systems", m_stage1), ("explicit_inverse", 1)):
        ms = eng_mod.bench_admm_blocks(p, K, 3, roof_iters, 3, m)
        modes[label] = {"kernel": "admm_persistent_kernel<K>", "blocks": m, "phases": 2 * m - 1,
                        "us_per_iteration": ms * 1e3, "achieved_gbs": alg_bytes_iter / (ms * 1e-3) / 1e9,
                        "frac": alg_bytes_iter / (ms * 1e-3) / 1e9 / hbm_peak}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_dmma_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("p") == p and tj.get("dram_bytes_per_pass"):
            traffic = tj["dram_bytes_per_pass"] * cg_passes
            traffic_src = ("imported (not measured in this run): DRAM bytes per matrix pass of the ncu --set full capture of "
                           "the same stream, " + tj.get("source", "profiles/r02_dmma_traffic.json"))
    # in the job: algorithmic bytes of every ADMM / matrix-free launch of the timed job over the sum of their
    # event-timed durations (per-launch CUDA events on each lane's stream; lanes share the GPU, so a launch's
    # duration includes waiting for SMs held by other lanes' kernels -- a lower bound on the kernels' own rate)
    ms_kernels = float(stats["ms_admm"]) + float(stats.get("ms_cg", 0.0))
    in_job = {"algorithmic_bytes": float(stats.get("admm_bytes", 0.0)), "ms_launches_summed": ms_kernels,
              "ms_matrix_free": float(stats.get("ms_cg", 0.0)), "ms_factored": float(stats["ms_admm"]),
              "matrix_free_calls": int(stats.get("n_cg_calls", 0)), "matrix_passes": int(stats.get("cg_passes", 0)),
              "achieved_gbs": (float(stats.get("admm_bytes", 0.0)) / (ms_kernels * 1e-3) / 1e9) if ms_kernels > 0 else None}
    if in_job["achieved_gbs"]:
        in_job["frac"] = in_job["achieved_gbs"] / hbm_peak
    if cg_ms:
        achieved = cg_flop_pass * cg_passes / (cg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": achieved / dgemm_tf,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": "batched_core2_kernel<MT=4> (fold-batched matrix-free spatpcaCore2 on the TMA-fed FP64 DMMA "
                              "stream: one launch = %d ADMM iterations of %d problems = %d matrix passes)"
                              % (roof_admm_iters, Mf, cg_passes),
                    "algorithmic_flop_per_launch": cg_flop_pass * cg_passes, "ms_per_launch": cg_ms,
                    "us_per_matrix_pass": cg_ms * 1e3 / max(cg_passes, 1),
                    "us_per_matrix_pass_per_problem": cg_ms * 1e3 / max(cg_passes, 1) / Mf,
                    "hbm_floor_us_per_single_problem_pass": 8.0 * p * p / (hbm_peak * 1e9) * 1e6,
                    "hbm_view": {"achieved_gbs": cg_bytes_pass * cg_passes / (cg_ms * 1e-3) / 1e9,
                                 "frac": cg_bytes_pass * cg_passes / (cg_ms * 1e-3) / 1e9 / hbm_peak},
                    "peak_source": "FP64 GEMM rate of this run (cuBLAS DGEMM 8192^3); MEASURED_PEAKS.json has no FP64 entry",
                    "hbm_peak": hbm_peak, "hbm_peak_source": peak_src,
                    "stream_alone": stream,
                    "single_problem_fma_kernel": ({"kernel": "cg_core2_kernel<K> (round-2 first cut, FMA stream; the fallback for odd n / K*M > 32)",
                                                   "us_per_matrix_pass": single_ms * 1e3 / max(single_passes, 1)} if single_ms else None),
                    "modes": modes, "in_job": in_job}
    else:
        ms_iter = modes["path_systems"]["us_per_iteration"] * 1e-3
        achieved = alg_bytes_iter / (ms_iter * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": None, "kernel": "admm_persistent_kernel<K>, %d-block factor (one launch = %d Core3 iterations)" % (m_path, roof_iters),
                    "algorithmic_bytes_per_launch": alg_bytes_iter * roof_iters, "ms_per_launch": ms_iter * roof_iters,
                    "peak_source": peak_src, "modes": modes, "in_job": in_job}
    # (3) the p^3 work (Omega build, the remaining SPD factorisations) against the FP64 GEMM rate of the same run
    inv_ms = eng_mod.bench_inverse(p, 2)
    fac_ms = eng_mod.bench_factor(p, 2, m_stage1)
    fac2_ms = eng_mod.bench_factor(p, 2, m_path)
    p3 = float(p) ** 3
    fac_flops = lambda m: p3 if m <= 1 else p3 / 3 + p3 / m + p3 / (m * m)
    fp64 = {"dgemm_8192_tflops": dgemm_tf, "inverse_ms": inv_ms, "inverse_tflops_p3": p3 / (inv_ms * 1e-3) / 1e12,
            "inverse_frac_of_dgemm": p3 / (inv_ms * 1e-3) / 1e12 / dgemm_tf,
            "factor_blocks": m_stage1, "factor_ms": fac_ms,
            "factor_frac_of_dgemm": fac_flops(m_stage1) / (fac_ms * 1e-3) / 1e12 / dgemm_tf,
            "path_factor_blocks": m_path, "path_factor_ms": fac2_ms,
            "path_factor_frac_of_dgemm": fac_flops(m_path) / (fac2_ms * 1e-3) / 1e12 / dgemm_tf,
            "job": {"p3_flops": float(stats.get("p3_flops", 0.0)), "n_factorisations": int(stats["n_inverses"]),
                    "ms_omega": float(stats["ms_omega"]), "ms_factorisations_summed": float(stats["ms_inverse"]),
                    "omega_note": "Omega build: FP64 Cholesky inverse of the rotated SPD block (potrf + GEMM-rate trtri + V'V, "
                                  "~1.1 p^3 FP64 flop) + one Newton-Schulz step and the final product on INT8 digit planes "
                                  "(78 + 15 + 28 plane products of 2 p^3 integer operations, upper strips); no single FP64 "
                                  "rate describes it, so only its stream time is reported"}}

    # ---- forced-iteration mode (SURVEY 8d mode ii: thr = 0, every ADMM call runs exactly maxit iterations): the
    # deterministic-denominator figure of BASELINE.json's "ADMM iters/s"; HBM floor = peak / (8 p^2 + 72 p K) ----
    forced = None
    if world == 1 and not args.no_forced:
        fm = args.forced_maxit
        cf = dict(case)
        e = sp.Engine(cf["x"].shape[0], cf["x"].shape[1], cf["Y"].shape[0], cf["M"], cf["K"], cf["tau1"], cf["tau2"],
                      cf["gamma"], fm, 0.0, cf["l2"])
        e.upload(cf["x"], cf["Y"], cf["nk"])
        cf["_engine"] = e
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D.prepare_sharded(e, D.Collective())
        D.run_sharded(e, sp.select_index, D.Collective(), cf["M"], len(cf["tau1"]), len(cf["tau2"]), len(cf["gamma"]),
                      cf["tau1"], cf["tau2"], cf["gamma"])
        torch.cuda.synchronize()
        dtf = time.perf_counter() - t0
        stf = e.stats()
        e.close()
        floor = hbm_peak * 1e9 / alg_bytes_iter
        forced = {"maxit": fm, "thr": 0.0, "wall_s": dtf, "admm_iters": int(stf["admm_iters"]),
                  "admm_calls": int(stf["n_admm_calls"]), "iters_per_s": stf["admm_iters"] / dtf,
                  "hbm_floor_iters_per_s": floor, "frac_of_hbm_floor": stf["admm_iters"] / dtf / floor,
                  "ms_admm_summed": float(stf["ms_admm"]) + float(stf.get("ms_cg", 0.0)),
                  "iters_per_s_kernel_time": (stf["admm_iters"] / ((float(stf["ms_admm"]) + float(stf.get("ms_cg", 0.0))) * 1e-3))
                  if stf["ms_admm"] + stf.get("ms_cg", 0.0) > 0 else None,
                  "note": "whole job incl. Omega and factorisations; with thr = 0 every call is expected to run maxit "
                          "iterations, so the cost model keeps the factored path"}

    # ---- CPU baseline on a bounded sample (oracle port, host cores) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import spatpca_ref as ref                     # checker / baseline only
        om = om_for_cpu                                            # sample input: Omega taken from the GPU run
        om[np.diag_indices(p)] += 1e-8
        dt, it, _ = cpu_sample(case, om)
        n_inv_ref = reference_inverse_count(case, res["index"])
        cpu = {"value": it / dt, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": ("fold 0 of stage 1 restricted to tau1[{0,5}] with Omega supplied: SVD start + Gram + 1 inv_sympd + "
                          f"{it} ADMM iterations + 2 losses at p={p} ({dt:.1f} s); the reference's full job has "
                          f"{n_inv_ref} inv_sympd calls (this repo's job: {int(stats['n_inverses'])} factorisations -- its ADMM "
                          "calls are solved matrix-free where the cost model allows)"),
               "sample_s": dt, "est_full_job_s": dt * n_inv_ref}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_job * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "wall_s": t_job, "admm_iters": iters_job, "admm_calls": int(stats["n_admm_calls"]),
        "selected": [res["selected_tau1"], res["selected_tau2"], res["selected_gamma"]],
        "phases_ms_rank0": {k: stats[k] for k in ("ms_omega", "ms_prologue", "ms_inverse", "ms_admm", "ms_cg", "ms_cv_loss", "ms_gamma")},
        "n_factorisations": int(stats["n_inverses"]), "matrix_free_calls": int(stats.get("n_cg_calls", 0)),
        "roofline": roofline, "fp64": fp64, "forced_iterations": forced, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches_job,
        "clocks": clocks,
    }
    emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
