#!/usr/bin/env python
"""bench.py — SQP iterations/sec (batched) of the collocated-NLP hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mobile_robot|cstr|kite] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm: the reference algorithm on the box's host cores

One "step" = one pmb_sqp_solve() of the whole batch: every instance runs SQPBase::solve to convergence (or max_iter) from
the same initial guess.  metric = sum over instances of sqp_info.iter / time.

  value : inputs resident in HBM when the timed region starts (bounds, x0 rows, guesses already on the device; the
          iterates are restored on the device by pmb_sqp_reset_guess) — K solves between two CUDA events.
  e2e   : the same solves through the C ABI with HOST buffers: every step copies x0 / guesses host->device
          (pmb_sqp_set_initial_conditions, pmb_sqp_set_primal, pmb_sqp_set_dual) and reads iterates and info back
          (pmb_sqp_get_primal, pmb_sqp_get_info) inside the timed region.
  roofline     : the dominant kernel of the step (the fused persistent sqp_solve), CUDA-event time per launch and its
                 phase split from SM cycle counters (pmb_sqp_set_profiling).
  kkt_kernel   : the materialising KKT kernel of the metric's second half (pmb_kkt_assemble_dev, box_admm.hpp:207-223),
                 B_KKT = 8 (N^2 + M N + (N+M)^2) bytes per instance, device-resident inputs larger than L2.
  cpu_baseline : the CPU restatement of the reference algorithm (oracle/, Eigen is not available so the reference itself
                 cannot be built) on a bounded sample of the same workload, all host cores — in two builds: "native" (-O3, AVX2 +
                 FMA contraction, libm: how a user would build the reference; the figure quoted) and "strict" (the bit-reproducible
                 parity oracle).
  arithmetic   : --arithmetic fast (default; KKT linear algebra on the fp64 tensor cores, csrc/pmb_qp_fast.hpp) or exact (every
                 fp64 operation in the oracle's order: bit-identical results).  The line always carries BOTH: the other mode's
                 throughput is under `other_arithmetic`, and `parity` holds the comparison of both with the oracle, next to the
                 control (the oracle against its own native build: what a change of rounding alone does to whole solves).
  configs      : the other BASELINE.json configurations (CSTR 4096, kite 1024 sharded over the ranks, robot batch sweep), short runs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from polympc_b200 import workloads as W  # noqa: E402

METRIC = "sqp_iterations_per_sec"
UNIT = "SQP iterations/s"
DEFAULT_BATCH = {"mobile_robot": 8192, "cstr": 4096, "kite": 1024}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def ncu_traffic(csv_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one committed `ncu --set full` capture (profiles/), bytes per launch"""
    try:
        tot, unit_mul = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        found = 0
        for line in open(os.path.join(ROOT, "profiles", csv_name)):
            f = line.strip().split(",")
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * unit_mul.get(f[2], 1.0); found += 1
        return tot if found == 2 else None
    except Exception:
        return None


def hbm_peak():
    p = measured_peaks()
    if p and p.get("hbm_gbs"):
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, n_inst: int, offset_seed: int = 0):
    fn = W.WORKLOADS[args.workload]
    w = fn(n_inst, sqp_max_iter=args.sqp_max_iter, ls_max_iter=args.ls_max_iter)
    return w


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the CPU restatement of the reference algorithm (oracle/), timed on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_solve_rate(args, n_inst: int, steps: int, warmup: int, keep=None, flavour="strict", trace=False):
    from oracle import pyoracle   # the one place bench.py executes oracle/: as the timed CPU arm
    orc = pyoracle.load_native() if flavour == "native" else pyoracle.load()
    cores = os.cpu_count() or 1
    pyoracle.load(); pyoracle.set_num_threads(cores)
    w = build_workload(args, max(args.batch, n_inst))        # the GPU arm's workload ...
    w.x0 = np.ascontiguousarray(w.x0[:n_inst])               # ... of which the CPU arm solves the first n_inst instances
    s = orc.sqp(w.name, n_inst)
    W.configure(s, w)
    total_it, total_t = 0, 0.0
    for k in range(warmup + steps):
        s.reset_guess()
        t0 = time.perf_counter()
        s.solve()
        dt = time.perf_counter() - t0
        if k >= warmup:
            total_it += int(s.info()["iter"].sum()); total_t += dt
    if keep is not None:
        keep["x"] = s.primal(); keep["lam"] = s.dual(); keep["info"] = s.info()
        if trace:
            keep["trace"] = s.trace(w.sqp_max_iter)
    s.close()
    return total_it / total_t, total_t / max(steps, 1), cores, total_it


def rel_inf_rows(a, b):
    return np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))


def compare_solves(a, b):
    """whole-solve comparison of two result sets (dicts with x, lam, info, trace) of the same instances — SURVEY.md §8d parity
    definitions: decision traces (ADMM trips, factorisations, BFGS branch, line-search trials, alpha per SQP iteration; iteration
    count; status), iterate parity on the instances whose traces are identical"""
    ia, ib = a["info"], b["info"]
    same = (ia["iter"] == ib["iter"]) & (ia["status"] == ib["status"])
    for k in ("qp_iter", "bfgs", "ls_trials", "qp_factor"):
        same &= (a["trace"][k] == b["trace"][k]).all(axis=1)
    same &= np.array([np.array_equal(u, v, equal_nan=True) for u, v in zip(a["trace"]["alpha"], b["trace"]["alpha"])])
    fin = np.isfinite(a["x"]).all(axis=1) & np.isfinite(b["x"]).all(axis=1)
    e = np.maximum(rel_inf_rows(a["x"], b["x"]), rel_inf_rows(a["lam"], b["lam"]))
    m = same & fin
    bit = ((a["x"] == b["x"]) | (np.isnan(a["x"]) & np.isnan(b["x"]))).all(axis=1) & ((a["lam"] == b["lam"]) | (np.isnan(a["lam"]) & np.isnan(b["lam"]))).all(axis=1)
    return {"instances": int(len(same)), "identical_status_pct": 100.0 * float((ia["status"] == ib["status"]).mean()),
            "identical_iteration_counts_pct": 100.0 * float((ia["iter"] == ib["iter"]).mean()),
            "identical_decision_traces_pct": 100.0 * float(same.mean()), "bit_identical_iterates_pct": 100.0 * float(bit.mean()),
            "max_rel_inf_on_identical_traces": float(e[m].max()) if m.any() else None,
            "median_rel_inf_on_identical_traces": float(np.median(e[m])) if m.any() else None,
            "within_1e-10_pct_of_identical_traces": 100.0 * float((e[m] <= 1e-10).mean()) if m.any() else None,
            "max_rel_inf_all_finite": float(e[fin].max()) if fin.any() else None}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_inst = min(args.batch, 8192) if args.workload == "mobile_robot" else args.cpu_sample      # about 1.6 s of CPU work per step
    rate, t_step, cores, _ = cpu_solve_rate(args, n_inst, args.steps, min(args.warmup, 1), flavour="native")
    sample = f"{n_inst} instances of the {args.workload} workload per step (same seeds/inputs as the GPU arm's first rows), solved to convergence"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.batch, max(1, args.gpus)),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of PolyMPC's algorithm (oracle/, native build: -O3, AVX2 + FMA contraction, libm): Eigen is absent, "
                                 "the reference itself cannot be built"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def config_dict(args, batch_per_gpu, n_gpus, cpu=False):
    sizes = {"mobile_robot": "NX=3,NU=2, Chebyshev order 6 x 2 segments (13 nodes), N=65, M=39",
             "cstr": "NX=4,NU=2, order 5 x 2 (11 nodes), N=66, M=44",
             "kite": "NX=13,NU=3, order 12 x 1 (13 nodes), N=208, M=169"}[args.workload]
    return {"workload": f"{args.workload} OCP batch={batch_per_gpu}{'' if cpu else ' per GPU'} random x0 sweep, fp64, SQP to convergence "
                        f"(max_iter {args.sqp_max_iter}, line search {args.ls_max_iter}), boxADMM + dense pivoted LDLT",
            "arithmetic": "reference order (CPU)" if cpu else args.arithmetic,
            "problem": sizes, "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n_gpus,
            "parallelism": "cpu-threads" if cpu else f"instances sharded over {n_gpus} GPU(s), no data-path collective",
            "l2": "inputs larger than L2 (H + A + state of one batch = %.0f MB)" % (batch_per_gpu * 8 * (65 * 65 + 39 * 65 + 6 * 65) / 1e6)
            if args.workload == "mobile_robot" else "inputs larger than L2"}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import polympc_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    api = polympc_b200.load()
    if api.device_count() < 1:
        raise SystemExit("bench.py: libpolympc_b200 sees no device")

    ARITH = {"exact": 0, "fast": 1}
    arith = ARITH[args.arithmetic]
    other = "exact" if args.arithmetic == "fast" else "fast"
    B = args.batch                                   # per GPU (weak scaling)
    w_all = build_workload(args, B * world)
    lo, hi = W.shard_bounds(B * world, world, rank)

    # shared Chebyshev tables: computed on rank 0, broadcast once at init, checked against the local tables (north_star)
    from polympc_b200 import distributed as PD
    dims = api.dims(w_all.name)
    PD.broadcast_tables(api, dims["P"], device=dev)

    stream = torch.cuda.Stream(device=dev)     # the engine launches on this stream, the events below are recorded on it
    torch.cuda.set_stream(stream)
    s = api.sqp(w_all.name, hi - lo, device=local_rank)
    s.set_stream(stream.cuda_stream)
    W.configure(s, w_all, lo, hi)
    s.set_arithmetic(arith)
    x0 = np.ascontiguousarray(w_all.x0[lo:hi])
    guess_x = np.zeros(dims["N"]); guess_l = np.zeros(dims["DUAL"])
    if w_all.x_guess is not None:
        guess_x[:dims["NX"] * dims["NN"]] = np.tile(w_all.x_guess, dims["NN"])
    if w_all.u_guess is not None:
        guess_x[dims["NX"] * dims["NN"]:dims["NX"] * dims["NN"] + dims["NU"] * dims["NN"]] = np.tile(w_all.u_guess, dims["NN"])
    def pinned(a):
        """copy of `a` in page-locked host memory (numpy view of a pinned torch tensor)"""
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(1, dtype=a.dtype)).dtype).pin_memory()
        v = t.numpy()
        v[...] = a
        pinned.keep.append(t)
        return v
    pinned.keep = []
    x0 = pinned(x0)
    guess_x_b = pinned(np.tile(guess_x, (hi - lo, 1)))
    guess_l_b = pinned(np.zeros((hi - lo, dims["DUAL"])))
    out_x = pinned(np.zeros((hi - lo, dims["N"])))
    out_info_t = torch.empty((hi - lo) * 12, dtype=torch.uint8).pin_memory()
    from polympc_b200.capi import SQP_INFO_DTYPE
    out_info = out_info_t.numpy().view(SQP_INFO_DTYPE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        return PD.max_over_ranks(v, device=dev)

    def sum_over_ranks(v: float) -> float:
        return PD.sum_over_ranks(v, device=dev)

    # ---- device-resident throughput ("value") -----------------------------------------------------------------------
    for _ in range(args.warmup):
        s.reset_guess(); s.solve()
    iters_per_solve = int(s.info()["iter"].sum())
    launches = 0
    kernel_ms_timed = 0.0
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        s.reset_guess()
        s.solve()
        launches += s.last_solve_launches()
        kernel_ms_timed += s.last_kernel_ms()     # CUDA events of the library around the fused launch, on the launching stream
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    total_iters = sum_over_ranks(float(iters_per_solve)) * args.steps
    value = total_iters / (ms_total * 1e-3)
    info = s.info()
    solved_frac = float((info["status"] == 0).mean())

    # ---- end to end through the C ABI with host buffers ("e2e") -----------------------------------------------------------
    h2d = x0.nbytes * 2 + guess_x_b.nbytes + guess_l_b.nbytes
    d2h = (hi - lo) * dims["N"] * 8 + (hi - lo) * 12
    for _ in range(max(1, args.warmup // 2)):
        s.set_initial_conditions(x0); s.set_primal(guess_x_b); s.set_dual(guess_l_b); s.solve(); s.primal(out_x); s.info(out_info)
    barrier()
    e0.record(stream)
    e2e_iters = 0
    for _ in range(args.steps):
        s.set_initial_conditions(x0)
        s.set_primal(guess_x_b)
        s.set_dual(guess_l_b)
        s.solve()
        xs = s.primal(out_x)
        e2e_iters += int(s.info(out_info)["iter"].sum())
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = sum_over_ranks(float(e2e_iters)) / (ms_e2e * 1e-3)

    # ---- the same device-resident measurement in the other arithmetic (always reported next to the headline) -----------------
    def device_rate(solver, steps, warm=2):
        for _ in range(warm):
            solver.reset_guess(); solver.solve()
        its = int(solver.info()["iter"].sum())
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(steps):
            solver.reset_guess(); solver.solve()
        a1.record(stream)
        barrier()
        ms = max_over_ranks(a0.elapsed_time(a1))
        return sum_over_ranks(float(its)) * steps / (ms * 1e-3), ms / steps

    other_arith = None
    try:
        s.set_arithmetic(ARITH[other])
        v_o, ms_o = device_rate(s, max(2, min(args.steps, 3)))
        other_arith = {"arithmetic": other, "value": v_o, "unit": UNIT, "ms_per_step": ms_o,
                       "note": "same workload, same handle, device-resident inputs, the other arithmetic of the KKT linear algebra"}
    except Exception as e:                                      # noqa: BLE001
        other_arith = {"arithmetic": other, "error": str(e)}
    s.set_arithmetic(arith)

    # ---- roofline of the dominant (only) kernel of the step: the fused persistent sqp_solve ---------------------------------------
    # CUDA-event time of the kernel alone (pmb_sqp_set_profiling brackets the launch) and its phase split (SM cycle counters)
    s.set_profiling(True)
    k_ms, k_n, phases = 0.0, 0, {}
    for _ in range(2):
        s.reset_guess(); s.solve()
        kt = s.kernel_times()
        k_ms += sum(v[0] for v in kt.values()); k_n += 1
        for k, v in s.phase_cycles().items():
            phases[k] = phases.get(k, 0) + v
    s.set_profiling(False)
    N, M = dims["N"], dims["M"]
    peak, peak_src = hbm_peak()
    kernel_ms = kernel_ms_timed / args.steps            # measured live over the timed region (no cycle counters in the kernel)
    kernel_ms_profiled = k_ms / k_n                    # the same launch with the phase counters on
    # algorithmic bytes per SQP iteration (compulsory state traffic, fp64):
    #   read+write x, lam, H, lag_grad_prev, step ; read lbx, ubx, lbg, ubg, d
    b_iter = 8 * (2 * (N * N + 4 * N + M) + 2 * N + dims["ND"])
    # flops per SQP iteration: QP = K^3/3 + trips (2 K^2 + 2 K) + floor(trips/10) 2 (2 M N + N^2); + BFGS and A^T lam
    it_n = max(1, phases.get("sqp_iterations", 1))
    trips = phases.get("admm_trips", 0) / it_n
    K = N + M
    flops_iter = K ** 3 / 3 + trips * (2 * K * K + 2 * K) + (trips / 10) * 2 * (2 * M * N + N * N) + 6 * N * N + 2 * M * N
    achieved = iters_per_solve * b_iter / (kernel_ms * 1e-3) / 1e9
    cyc_tot = sum(phases.get(k, 0) for k in ("linearise", "qp", "step")) or 1
    traffic_csv = "r2_sqp_solve_%s_ncu_raw.csv" % args.arithmetic
    roofline = {"kernel": ("sqp_solve_fast" if arith else "sqp_solve") + " (fused persistent kernel: linearise + boxADMM/LDLT + line search, one CTA per instance)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(traffic_csv) if (args.workload == "mobile_robot" and B == 8192) else None,
                "traffic_source": "profiles/" + traffic_csv + " (one ncu --set full capture of this launch at batch 8192, taken at the commit "
                                  "named in profiles/README.md; null when the file is absent)",
                "peak_source": peak_src, "avg_launch_ms": kernel_ms, "avg_launch_ms_with_phase_counters": kernel_ms_profiled, "bytes_per_sqp_iteration": b_iter,
                "sqp_iterations_per_launch": iters_per_solve,
                "kernel_share_of_step": kernel_ms / (ms_total / args.steps),
                "phase_share": {k: phases.get(k, 0) / cyc_tot for k in ("linearise", "qp", "step")},
                "qp_phase_share": {k: phases.get(k, 0) / max(1, phases.get("qp", 1)) for k in
                                   ("qp_pivot", "qp_gather", "qp_factor", "qp_solve", "qp_update", "qp_resid")},
                "cycles_per_sqp_iteration": {k: int(v / it_n) for k, v in phases.items() if k not in ("sqp_iterations", "admm_trips")},
                "admm_trips_per_iteration": trips,
                "fp64": {"achieved_tflops": iters_per_solve * flops_iter / (kernel_ms * 1e-3) / 1e12, "peak_tflops": 37.07,
                         "peak_source": "measured on this pool's B200 with tools/ubench/lat.cu (fp64 FMA, all SMs)",
                         "flops_per_sqp_iteration": flops_iter},
                "note": "fused design: K is built, factored and used in shared memory and never written to HBM; per-iteration state "
                        "stays in L2.  The kernel is bound by instruction issue / latency (barriers, instruction fetch, dependent fp64 "
                        "chains), not by HBM or fp64 throughput; the HBM-bound materialising KKT kernel is reported under kkt_kernel"}

    # ---- the materialising KKT kernel (a17), device-resident, B_KKT bytes per instance ----------------------------------------------
    kkt = None
    try:
        nb = hi - lo
        H_d = torch.randn(nb, N * N, device=dev, dtype=torch.float64)
        A_d = torch.randn(nb, M * N, device=dev, dtype=torch.float64)
        rb_d = torch.rand(nb, N, device=dev, dtype=torch.float64) + 0.1
        ri_d = torch.rand(nb, M, device=dev, dtype=torch.float64) + 0.1
        K_d = torch.empty(nb, (N + M) * (N + M), device=dev, dtype=torch.float64)
        fn = api._fn("kkt_assemble_dev")
        def kkt_launch():
            rc = fn(N, M, nb, H_d.data_ptr(), A_d.data_ptr(), rb_d.data_ptr(), ri_d.data_ptr(), 1e-6, K_d.data_ptr(), stream.cuda_stream)
            assert rc == 0
        for _ in range(3):
            kkt_launch()
        torch.cuda.synchronize()
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            kkt_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        bk = 8 * (N * N + M * N + (N + M) * (N + M))
        gbs = nb * bk / (kms * 1e-3) / 1e9
        kkt = {"kernel": "kkt_assemble_dense", "bytes_per_instance": bk, "batch": nb, "avg_launch_ms": kms, "achieved": gbs, "peak": peak,
               "unit": "GB/s", "frac": gbs / peak, "working_set_mb": nb * bk / 1e6,
               "traffic": ncu_traffic("r2_kkt_assemble_ncu_raw.csv") if (args.workload == "mobile_robot" and nb == 8192) else None}
        launches_kkt = reps
        del H_d, A_d, K_d
    except Exception as e:   # never let the auxiliary measurement kill the bench line
        kkt = {"error": str(e)}

    # ---- CPU baseline + parity (rank 0, N == 1 only) -----------------------------------------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample, hi - lo)
        strict, native = {}, {}
        cpu_solve_rate(args, ns, 1, 0, keep=strict, flavour="strict", trace=True)      # results for the parity object (untimed)
        cpu_solve_rate(args, ns, 1, 0, keep=native, flavour="native", trace=True)
        nb_cpu = (hi - lo) if args.workload == "mobile_robot" else ns                  # timed: about 10 s of CPU work
        rate_n, t_n, cores, _ = cpu_solve_rate(args, nb_cpu, 5, 1, flavour="native")
        rate_s, t_s, _, _ = cpu_solve_rate(args, nb_cpu, 2, 1, flavour="strict")
        cpu = {"value": rate_n, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"all {nb_cpu} instances of the same workload, 5 solves to convergence after 1 warm-up ({t_n:.2f} s each), all host cores",
               "build": "native: -O3 -march=x86-64-v3 -ffp-contract=fast, libm (oracle/Makefile)",
               "strict_build": {"value": rate_s, "seconds": t_s, "build": "-O3 -march=x86-64-v3 -ffp-contract=off, deterministic elementary "
                                "functions: the bit-reproducible parity oracle"},
               "note": "CPU restatement of PolyMPC's algorithm (oracle/); Eigen is absent so the reference itself cannot be built"}
        # the CPU legs solved the first `ns` instances of the same workload: the GPU's results for them, in both arithmetics
        gpu = {}
        sp = api.sqp(w_all.name, ns, device=local_rank); sp.set_stream(stream.cuda_stream); W.configure(sp, w_all, lo, lo + ns); sp.set_trace(True)
        for name in ("exact", "fast"):
            try:
                sp.set_arithmetic(ARITH[name]); sp.reset_guess(); sp.solve()
                gpu[name] = {"x": sp.primal(), "lam": sp.dual(), "info": sp.info(), "trace": sp.trace(w_all.sqp_max_iter)}
            except Exception as e:                              # noqa: BLE001
                gpu[name] = None
        sp.close()
        # stage-wise: ONE SQP iteration from the same state (exact Hessian, QP, line search, step), fast vs exact on the GPU
        one = None
        try:
            w1 = build_workload(args, ns); w1.sqp_max_iter = 1
            r1 = {}
            for name in ("exact", "fast"):
                s1 = api.sqp(w1.name, ns, device=local_rank); s1.set_stream(stream.cuda_stream); W.configure(s1, w1); s1.set_arithmetic(ARITH[name]); s1.solve()
                r1[name] = (s1.primal(), s1.dual(), s1.info()); s1.close()
            e1 = np.maximum(rel_inf_rows(r1["fast"][0], r1["exact"][0]), rel_inf_rows(r1["fast"][1], r1["exact"][1]))
            one = {"instances": int(ns), "max_rel_inf": float(e1.max()), "tolerance": 1e-10,
                   "identical_admm_trip_counts_pct": 100.0 * float((r1["fast"][2]["qp_solver_iter"] == r1["exact"][2]["qp_solver_iter"]).mean())}
        except Exception as e:                                  # noqa: BLE001
            one = {"error": str(e)}
        parity = {"tolerance": 1e-10, "against": "CPU restatement (oracle/, strict build) on the same inputs: the first %d instances" % ns,
                  "exact_vs_oracle": compare_solves(gpu["exact"], strict) if gpu.get("exact") else None,
                  "fast_vs_oracle": compare_solves(gpu["fast"], strict) if gpu.get("fast") else None,
                  "fast_one_sqp_iteration_vs_exact": one,
                  "control_native_oracle_vs_oracle": compare_solves(native, strict),
                  "note": "exact arithmetic is bit-identical to the oracle.  Fast arithmetic rounds differently: one SQP iteration from the "
                          "same state agrees to 1e-12; over whole solves rounding is amplified by the iteration itself (Armijo and "
                          "termination tests at round-off level) — the control line shows the oracle against its own native build, "
                          "i.e. what a change of rounding ALONE does; fast-vs-oracle should be read against it"}

    # ---- two batches in flight (secondary figure, never the headline) ---------------------------------------------------------
    # A persistent kernel ends with a tail: the 0.4 % of instances that run all 100 SQP iterations keep a few CTAs busy for
    # ~50 ms while the other SMs idle.  A caller that owns several independent batches hides that tail by
    # keeping two solves in flight on two streams: CTAs of the second launch become resident as CTAs of the first retire.
    # Same work per step as `value` (every step solves the whole batch from the same guess); steps alternate between two
    # solver handles through pmb_sqp_solve_async / pmb_sqp_wait.
    pipelined = None
    if world == 1 and not args.no_pipelined:
        try:
            stream2 = torch.cuda.Stream(device=dev)
            s2 = api.sqp(w_all.name, hi - lo, device=local_rank)
            s2.set_stream(stream2.cuda_stream)
            W.configure(s2, w_all, lo, hi)
            s2.set_arithmetic(arith)
            s2.solve(); s2.reset_guess(); s2.solve()
            steps2 = max(2, args.steps + (args.steps % 2))
            barrier()
            p0, p1, pj = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
            p0.record(stream); stream2.wait_event(p0)
            for k in range(steps2):                 # pmb_sqp_solve_async / pmb_sqp_wait: one host thread, two handles
                sv = (s, s2)[k % 2]
                sv.wait(); sv.reset_guess(); sv.solve_async()
            s.wait(); s2.wait()
            pj.record(stream2); stream.wait_event(pj); p1.record(stream)
            barrier()
            ms_p = p0.elapsed_time(p1)
            pipelined = {"value": iters_per_solve * steps2 / (ms_p * 1e-3), "unit": UNIT, "batches_in_flight": 2, "steps": steps2,
                         "ms_per_step": ms_p / steps2,
                         "note": "two independent batches in flight on two streams hide the straggler tail of the persistent "
                                 "kernel; `value` above is one batch at a time"}
            s2.close()
        except Exception as e:                                  # noqa: BLE001 — a secondary figure must not break the line
            pipelined = {"error": str(e)}

    # ---- the other BASELINE.json configurations, short runs (every rank takes part: max-over-ranks timing) ---------------------
    configs = None
    if not args.no_configs:
        configs = {}

        def quick(workload, batch_per_rank, arithmetic, steps=2, options=None, **kw):
            """it/s of `workload` with batch_per_rank instances on every rank (rank r solves block r of the sweep)"""
            try:
                wq = W.WORKLOADS[workload](batch_per_rank * world, **kw)
                sq = api.sqp(wq.name, batch_per_rank, device=local_rank)
                sq.set_stream(stream.cuda_stream)
                W.configure(sq, wq, rank * batch_per_rank, (rank + 1) * batch_per_rank)
                sq.set_arithmetic(ARITH[arithmetic])
                for name, val in (options or {}).items():      # solver options of the reference's other operators (capi.Sqp setters)
                    getattr(sq, name)(*val)
                v, ms = device_rate(sq, steps, warm=2)
                solved = sum_over_ranks(float((sq.info()["status"] == 0).sum())) / (batch_per_rank * world)
                sq.close()
                return {"value": v, "unit": UNIT, "ms_per_step": ms, "batch_per_gpu": batch_per_rank, "arithmetic": arithmetic, "solved_fraction": solved}
            except Exception as e:                              # noqa: BLE001
                return {"error": str(e)}

        configs["cstr_batch4096"] = {a: quick("cstr", 4096, a) for a in ("fast", "exact")}                      # BASELINE config 3
        kb = max(1, 1024 // world)
        configs["kite_12x1_batch1024_sharded"] = {"exact": quick("kite", kb, "exact", steps=1),                   # config 4: 1024 / N per GPU
                                                  "fast": quick("kite", kb, "fast", steps=1),
                                                  "note": "kite is quoted in exact arithmetic: fast arithmetic loses 2e-7 per SQP iteration on its "
                                                          "377 x 377 KKT systems (tests/test_gpu_fast.py)"}
        configs["mobile_robot_batch_sweep"] = {str(b): quick("mobile_robot", b, args.arithmetic) for b in (256, 1024, 4096, 16384, 65536)}   # config 5
        configs["mobile_robot_reference_test_settings"] = quick("mobile_robot", B, args.arithmetic, sqp_max_iter=10, ls_max_iter=10)
        # the other operators behind the same API, robot batch B: the solver of the reference's
        # control tests (block BFGS) with RuizEquilibration<SPARSE> around every QP, and ADMM<> as the QP solver.  (The filter line
        # search is left out here: its filter lives across solves, so repeated timed solves are not the same work.)
        configs["mobile_robot_block_bfgs_ruiz"] = quick("mobile_robot", B, args.arithmetic, options={
            "set_hessian_update": (1,), "set_preconditioner": (2,)})
        configs["mobile_robot_osqp_style_admm"] = quick("mobile_robot", B, "exact", options={"set_qp_solver": (1,)})
        configs["note"] = "per-GPU batch; under --gpus N every rank solves its own block of the sweep (weak scaling) except the kite, which is sharded"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, B, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "arithmetic": args.arithmetic, "other_arithmetic": other_arith,
            "roofline": roofline, "kkt_kernel": kkt, "cpu_baseline": cpu, "parity": parity, "pipelined": pipelined, "configs": configs,
            "schedule": "work queue in longest-processing-time-first order from the previous solve's iteration counts (pmb_sqp_set_schedule; "
                        "the warm-up solves provide the history, results do not depend on the order)",
            "sqp_iterations_per_step": total_iters / args.steps, "solved_fraction": solved_frac,
            "mean_sqp_iter_per_instance": iters_per_solve / (hi - lo),
        }
        print(json.dumps(line))
    s.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. NCCL's version banner) is
    # sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(real_stdout, "w")
    _print = print

    def emit(line):
        _print(line, file=json_out, flush=True)

    globals()["print"] = lambda *a, **k: emit(a[0]) if (len(a) == 1 and isinstance(a[0], str) and a[0].startswith("{")) else _print(*a, **k)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mobile_robot", choices=sorted(W.WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU")
    ap.add_argument("--sqp-max-iter", type=int, default=100)
    ap.add_argument("--ls-max-iter", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=None, help="instances solved by the CPU arm per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the two-batches-in-flight secondary measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE.json configurations")
    ap.add_argument("--arithmetic", default="fast", choices=["fast", "exact"], help="arithmetic of the KKT linear algebra (pmb_sqp_set_arithmetic)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.workload]
    if args.cpu_sample is None:
        args.cpu_sample = {"mobile_robot": 8192, "cstr": 4096, "kite": 64}[args.workload]    # the parity object covers the whole default batch
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
