#!/usr/bin/env python
"""bench.py - candidate-points/s through fused GP posterior-predict + EI (BASELINE.json metric).

Workload (BASELINE config 3): d=16, N_train=4096, Matern-2.5 (l=0.7, alpha=1e-6, normalize_y),
EI xi=0.01, M = 2^20 uniform candidates per GPU per step, fp64.  One "step" = one pass of the hot
path over one candidate batch: K* build -> V = L^-1 K*^T -> (mu, sigma) -> EI -> argmin + top-10.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Keys beyond the base contract:
  value     candidates already resident in HBM (device entry point of the C ABI), whole job
  e2e       same metric through the public Python API with HOST buffers: pinned host -> device
            copy of the candidates and device -> host read of the argmin/top-k records per step
  roofline  dominant kernel (predict_acq_kernel) vs the roof that binds it
  cpu_baseline  the reference's CPU path (sklearn GaussianProcessRegressor.predict + the restated
            closure, oracle/gp_oracle.py) on a bounded sample, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, N_TRAIN, LS, ALPHA, XI, KSEEDS = 16, 4096, 0.7, 1e-6, 0.01, 10
M_PER_GPU = 1 << 20
N_CAND_BUFFERS = 3  # rotated so that no step re-reads candidates from L2
NCU_DRAM_BYTES_PER_LAUNCH = 569.44e9 + 34.40e9  # from the committed ncu capture (see traffic_source)


def flops_per_candidate(n, d, n_gps=1):
    """F = N^2 + N(3d + 18) per GP."""
    return n_gps * (n * n + n * (3 * d + 18))


def hbm_model_bytes_per_candidate(n, d, tile, s=8):
    """SURVEY.md 8(d) 'TRSM-bound HBM' model: L^-1 re-streamed once per tile of T candidates."""
    return s * n * (n + 1) / (2 * tile) + s * n * (d + 1) / tile + s * d + s


def make_problem(seed=0):
    rs = np.random.RandomState(seed)
    X = rs.uniform(size=(N_TRAIN, D))
    y = np.sin(X.sum(1)) + 0.1 * rs.randn(N_TRAIN)
    return X, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(X, y, m_sample, reps=1):
    """The reference's CPU path: sklearn GaussianProcessRegressor.predict(return_std) (the
    unmodified dependency the reference calls, R/bayes_opt/acquisition.py:216) + the closure
    restated in oracle/gp_oracle.py, chunked at 2^14 rows."""
    import warnings

    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern

    from oracle import gp_oracle as O

    gp = GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=LS), alpha=ALPHA, normalize_y=True,
                                  optimizer=None).fit(X, y)
    y_max = float(y.max())
    xt = np.random.RandomState(1).uniform(size=(m_sample, D))

    def closure(x):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = []
            for i in range(0, len(x), 1 << 14):
                mu, sd = gp.predict(x[i:i + (1 << 14)], return_std=True)
                out.append(-1 * O.base_acq(O.ACQ_EI, mu, sd, xi=XI, y_max=y_max))
        return np.concatenate(out)

    closure(xt[:1024])  # warm-up
    best = np.inf
    for _ in range(reps):
        t0 = time.perf_counter()
        ys = closure(xt)
        int(np.argmin(ys)), np.argsort(ys)[:KSEEDS]
        best = min(best, time.perf_counter() - t0)
    return m_sample / best, best


def base_line(args, world):
    return {
        "metric": "candidate-pts/s GP-predict+EI @ N_train=4096,d=16",
        "unit": "candidates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: d=16, N_train=4096, Matern-2.5 l=0.7, EI xi=0.01, "
                               f"M=2^20 candidates per GPU per step, argmin+top-{KSEEDS}",
                   "n_train": N_TRAIN, "d": D, "candidates_per_gpu_per_step": M_PER_GPU,
                   "l2": f"{N_CAND_BUFFERS} candidate buffers rotated (128 MiB each) + 67 MB factor + "
                         "0.6 GB K* scratch per step: working set > 126 MB L2",
                   "parallelism": f"candidates sharded x{world}, model replicated"},
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    X, y = make_problem()
    m_sample = 1 << 13
    import threadpoolctl

    cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host thread
    threadpoolctl.threadpool_limits(limits=cores)
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_rate(X, y, 1024)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, _ = cpu_reference_rate(X, y, m_sample)
        rates.append(r)
    wall = time.perf_counter() - t0
    v = float(np.median(rates))
    line = base_line(args, world)
    line.update({
        "impl": "reference", "value": v, "ms_per_step": 1e3 * m_sample / v,
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{m_sample} candidates/step x {args.steps} steps, sklearn "
                                   "GaussianProcessRegressor.predict + restated EI closure, chunks of 2^14",
                         "blas": [i.get("internal_api") + ":" + str(i.get("num_threads"))
                                  for i in threadpoolctl.threadpool_info()]},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    })
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun's OMP notice) write to fd 1; the contract is ONE
    JSON line on stdout.  Point fd 1 at stderr for the whole run and keep the real stdout aside."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-mode", action="store_true", help="skip the secondary fp32-mode line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import bayesianoptimization_b200 as bo
    from bayesianoptimization_b200 import _lib as B
    from sklearn.gaussian_process.kernels import Matern

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = B.lib()

    # ---- model: every rank factorises the same data (deterministic, no traffic) -------------
    X, y = make_problem()
    gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=LS), alpha=ALPHA,
                                         normalize_y=True, optimizer=None, device=local_rank)
    t0 = time.perf_counter()
    gp.fit(X, y)
    fit_s = time.perf_counter() - t0
    ei = bo.ExpectedImprovement(xi=XI)
    ei.y_max = float(y.max())
    acq = ei._get_acq(gp=gp)

    # ---- candidates: host MT19937 (pinned) ; device copies for the HBM-resident leg ----------
    m = M_PER_GPU
    host_bufs, dev_bufs = [], []
    for b in range(N_CAND_BUFFERS):
        rs = np.random.RandomState(1000 + 17 * rank + b)
        t = torch.from_numpy(rs.uniform(size=(m, D))).pin_memory()
        host_bufs.append(t)
        dev_bufs.append(t.to(dev))
    sel = torch.zeros((KSEEDS + 1, 2), dtype=torch.int64, device=dev)
    sel_host = torch.zeros((KSEEDS + 1, 2), dtype=torch.int64).pin_memory()
    stream = torch.cuda.current_stream()
    index_base = rank * m

    gathered = torch.zeros((world, KSEEDS + 1, 2), dtype=torch.int64, device=dev) if world > 1 else None

    def exchange():
        # the path's ONE exchange step: all_gather of the (argmin, top-k) records, 176 B per rank
        if world > 1:
            dist.all_gather_into_tensor(gathered, sel)

    def step_device(i):
        B.check(L.b200bo_acq_eval_dev(C.byref(acq.spec), dev_bufs[i % N_CAND_BUFFERS].data_ptr(), m, None,
                                      None, None, KSEEDS, sel.data_ptr(), index_base, stream.cuda_stream))
        exchange()

    host_np = [t.numpy() for t in host_bufs]  # numpy views of the pinned buffers

    def step_e2e(i):
        # the call a user of the package makes (AcquisitionFunction._random_sample_minimize does exactly
        # this): host candidates in, (argmin, value, seed indices) out.  Inside: finite check, H2D copy
        # of the batch from pinned host memory, fused kernel + selection, D2H of the records.
        idx, val, top = acq.argmin_topk(host_np[i % N_CAND_BUFFERS], KSEEDS)
        if world > 1:  # the exchange step, from host-side records
            rec = torch.full((KSEEDS + 1, 2), -1, dtype=torch.int64)
            rec[0, 0] = int(np.float64(val).view(np.int64))
            rec[0, 1] = index_base + idx
            rec[1:1 + len(top), 1] = torch.from_numpy(index_base + np.asarray(top, dtype=np.int64))
            dist.all_gather_into_tensor(gathered, rec.to(dev))
        return idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        kernel_ms = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.b200bo_launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev0.record()
        for i in range(steps):
            fn(warmup + i)
            ms = C.c_float()
            B.check(L.b200bo_last_kernel_ms(C.byref(ms)))  # CUDA events on the launching stream
            kernel_ms.append(ms.value)
        ev1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        total_ms = ev0.elapsed_time(ev1)
        launches = L.b200bo_launch_count() - launches0
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, kernel_ms, launches, clocks

    total_ms, kernel_ms, launches, clocks = timed(step_device, args.steps, args.warmup)
    value = world * m * args.steps / (total_ms * 1e-3)

    # merge of the gathered per-rank records (argmin + seeds) with (value, index) ordering
    stream.synchronize()
    if world > 1:
        from bayesianoptimization_b200.sharding import merge_selection

        allr = gathered.cpu().numpy()
        best_idx, best_val, _ = merge_selection(allr.view(np.float64)[:, :, 0], allr[:, :, 1], KSEEDS)
    else:
        sel_np = sel.cpu().numpy()
        best_idx, best_val = int(sel_np[0, 1]), float(sel_np.view(np.float64)[0, 0])

    e2e_ms, _, _, _ = timed(step_e2e, args.steps, 1)
    e2e_value = world * m * args.steps / (e2e_ms * 1e-3)

    # ---- secondary line: fp32 mode (N^2 term as 3xTF32 on tcgen05), same workload ------------
    fp32 = None
    if not args.no_fp32_mode:
        gp32 = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=LS), alpha=ALPHA,
                                               normalize_y=True, optimizer=None, device=local_rank,
                                               precision="fp32")
        gp32.fit(X, y)
        acq32 = ei._get_acq(gp=gp32)
        sel32 = torch.zeros((KSEEDS + 1, 2), dtype=torch.int64, device=dev)

        def step_fp32(i):
            B.check(L.b200bo_acq_eval_dev(C.byref(acq32.spec), dev_bufs[i % N_CAND_BUFFERS].data_ptr(), m, None,
                                          None, None, KSEEDS, sel32.data_ptr(), index_base, stream.cuda_stream))

        t32_ms, k32_ms, _, clocks32 = timed(step_fp32, args.steps, args.warmup)
        stream.synchronize()
        s32 = sel32.cpu().numpy()
        fp32 = {"total_ms": t32_ms, "kernel_ms": float(np.mean(k32_ms)), "clocks": clocks32,
                "argmin_index": int(s32[0, 1]), "argmin_value": float(s32.view(np.float64)[0, 0])}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        k_ms = float(np.mean(kernel_ms))
        flops = flops_per_candidate(N_TRAIN, D) * m
        fp64_peak = 64 * 2 * 148 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12  # DFMA/clk/SM nominal
        peak_src = ("nominal 64 FMA/clk/SM x 148 SM x sm_max_mhz (fp64 has no tcgen05 path; not in "
                    "MEASURED_PEAKS.json)")
        try:
            mb = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_microbench.json")))
            fp64_peak = max(v for k, v in mb.items() if k.startswith("dmma"))
            peak_src = ("measured on this pool: best mma.sync f64 (DMMA) rate of tools/microbench.cu, "
                        "profiles/r01_fp64_microbench.json (nominal 37.2; MEASURED_PEAKS.json has no fp64 entry)")
        except Exception:
            pass
        ach_tf = flops / (k_ms * 1e-3) / 1e12
        hbm_bytes = hbm_model_bytes_per_candidate(N_TRAIN, D, 128) * m
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = base_line(args, world)
        line.update({
            "value": value, "ms_per_step": total_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": m * D * 8,
                    "d2h_bytes_per_step": (KSEEDS + 1) * 16, "ms_per_step": e2e_ms / args.steps,
                    "api": "FusedAcquisition.argmin_topk(host ndarray, k) - the package's public call: finite check, "
                           "H2D of the batch from pinned host memory, fused kernel + selection, D2H of the records"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "tensor", "pipe": "fp64 tensor path: mma.sync m8n8k4 f64 (SASS DMMA; ncu "
                "sm__pipe_tensor_subpipe_dmma) - tcgen05.mma has no f64 kind",
                "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, ncu "
                                  "--set full capture profiles/r01_final_predict_acq_dmma_ncu_summary.txt (K* scratch "
                                  "streams through HBM: 0.6 GB working set > L2)",
                "kernel": "predict_acq_kernel", "kernel_ms": k_ms,
                "peak_source": peak_src,
                "algorithmic_flops_per_candidate": flops_per_candidate(N_TRAIN, D),
                "hbm_model": {"achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "tile_T": 128,
                              "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s"},
            },
            "fit_seconds_fixed_theta": fit_s,
            "result": {"argmin_index": best_idx, "argmin_value": best_val},
        })
        if fp32 is not None:
            tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0
            v32 = world * m * args.steps / (fp32["total_ms"] * 1e-3)
            executed = 3.0 * N_TRAIN * N_TRAIN * m / (fp32["kernel_ms"] * 1e-3) / 1e12  # 3 TF32 products / MAC pair
            line["fp32_mode"] = {
                "value": v32, "unit": "candidates/s", "dtype": "tf32x3 (fp32 accumulate in TMEM); K*, mean, "
                "epilogue fp64", "tolerance": "1e-3 rel on the predictive variance (+1e-4 s_y^2)",
                "kernel": "predict_acq_tc_kernel (tcgen05.mma kind::tf32, bulk-copy producer, TMEM epilogue)",
                "kernel_ms": fp32["kernel_ms"], "clocks": fp32["clocks"],
                "roofline": {"bound": "tensor", "achieved": executed, "peak": tf32_peak, "unit": "TFLOP/s",
                             "frac": executed / tf32_peak,
                             "note": "executed TF32 tensor flops (3 products per useful multiply-add); dense TF32 "
                                     "peak taken as half the measured bf16 peak of MEASURED_PEAKS.json",
                             "useful_tflops": flops / (fp32["kernel_ms"] * 1e-3) / 1e12},
                "argmin_matches_fp64": fp32["argmin_index"] == best_idx,
                "argmin_value": fp32["argmin_value"],
            }
        if world == 1 and not args.no_cpu_baseline:
            v, secs = cpu_reference_rate(X, y, 1 << 15)
            line["cpu_baseline"] = {
                "value": v, "unit": "candidates/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"32768 of the same candidates ({secs:.1f} s): sklearn GaussianProcessRegressor."
                          "predict + restated EI closure (oracle/gp_oracle.py), chunks of 2^14"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
