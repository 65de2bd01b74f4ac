#!/usr/bin/env python
"""Benchmark of the hot path: converged ascent-NLP solves per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch: every rank solves its own 65,536-instance
dispersion batch (BASELINE.json configs[3]: thrust, Isp, initial mass, angular-acceleration
limit, target perilune/apolune; nt=200, NODES=2) to a scaled KKT error of 1e-10, then (N>1) one
NCCL allgather of the per-problem results.  Weak scaling: per-GPU work is fixed.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (parameters already in
HBM, results left in HBM); `e2e` is the same metric through the public host API
(`AscentSolver.solve_rows` on pinned host tensors = lmato_solve_batch_host: H2D of the
parameters, solve, D2H of tf / final mass / status / iterations / KKT error and the full
[10, nt, B] trajectory block).
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per mesh stage per IPM iteration
FLOP_PER_STAGE = 2147.0
BYTE_PER_STAGE = 336.0
METRIC = "converged ascent-NLP solves/sec at batch 64K"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE ascent_ipm_kernel launch, from the `ncu --set full`
# capture summarised in profiles/r01_ncu_ascent_ipm_kernel_metrics.csv; keyed by (dcost on, batch, nt) and
# only reported when the run matches that workload, otherwise null.
NCU_TRAFFIC_BYTES = {(True, 65536, 200): 217.488e9 + 123.538e9}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end leg (first touch) live on the NUMA node the GPU's PCIe root hangs on.  Multi-rank runs only:
    eight ranks each stream 1 GB of results per step into host memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region, every 200 ms, through NVML
    (nvidia_ml_py) from a thread.  (An `nvidia-smi -lms 200` subprocess was measured to cost ~17 %
    of a 140 ms step here, so the same counters are read in-process instead; nvidia-smi is the
    fallback.)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, reasons-bitmask)
        self.stop_flag = threading.Event()
        self.thread = None
        self.mode = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.mode = "nvml"
            # prime every query once (the first NVML call of each kind is slow)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            self.stop_flag.wait(float(os.environ.get("LMATO_CLOCK_PERIOD_S", "0.2")))

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nv
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [s[0] for s in self.samples]
            mx = [s[1] for s in self.samples]
            reasons = sorted({n for n, bit in names.items() for s in self.samples if s[2] & bit})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.mode != "smi" or self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on all host cores.  GEKKO/apm/IPOPT are not
    installable here (no network, not in /opt/wheelhouse), so this is the oracle port."""
    if rank != 0:
        return
    import numpy as np
    import torch
    import lunar_module_ascent_trajectory_optimiser_b200 as lm
    from oracle.cpu_baseline import OraclePool
    pool = OraclePool()
    per_step = pool.cores                      # bounded sample: one problem per core per step
    rows = lm.dispersed_params(max(per_step * (args.steps + args.warmup), 1), seed=11).rows().numpy()
    opts = lm.SolverOptions()
    off = 0
    for _ in range(args.warmup):
        pool.solve(rows[:, off:off + per_step], args.nt, opts.tol, opts.obj_scale); off += per_step
    t0 = time.perf_counter()
    nconv = 0
    iters = []
    for _ in range(args.steps):
        tf, st, it, _w = pool.solve(rows[:, off:off + per_step], args.nt, opts.tol, opts.obj_scale)
        off += per_step
        nconv += int((st == 0).sum()); iters += it.tolist()
    wall = time.perf_counter() - t0
    pool.close()
    val = nconv / wall
    sample = f"{per_step} problems per step (one per host core) of the seed-11 cfg4 dispersions, nt={args.nt}, tol={opts.tol:g}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(args.steps, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg4: elliptical ascent, nt={args.nt}, NODES=2, 6-parameter dispersions (seed 11)",
                       "sample_per_step": per_step, "tol": opts.tol},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": pool.cores, "kind": "port", "sample": sample,
                             "note": "oracle restatement (numpy/scipy sparse IPM), not GEKKO/IPOPT: those are absent from this image",
                             "mean_iterations": float(np.mean(iters)) if iters else None},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--nt", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traj", action="store_true", help="do not materialise the [10,nt,B] trajectory block")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3                      # timing rules: W >= 3
    import torch
    import torch.distributed as dist
    import lunar_module_ascent_trajectory_optimiser_b200 as lm
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, nt, N = args.batch, args.nt, args.nt - 1
    traj = not args.no_traj
    opts = lm.SolverOptions()
    solver = lm.AscentSolver(lm.Mesh(nt=nt), opts, device=dev)
    # Successive steps get DIFFERENT batches (three draws of the same dispersion, cycled), so nothing
    # a step computes -- including the warm start's reference solve, which starts from the previous
    # call's reference -- can be a replay of the step before.
    NBATCH = 3
    rows_hosts = [lm.dispersed_params(B, seed=11 + rank + 1000 * j).rows(B).pin_memory() for j in range(NBATCH)]
    rows_devs = [r.to(dev) for r in rows_hosts]
    rows_host, rows_dev = rows_hosts[0], rows_devs[0]
    step_no = [0]
    gathered = torch.empty((world * B, 4), dtype=torch.float64, device=dev) if world > 1 else None

    out_dev = solver.alloc_outputs(B, traj, on_device=True)      # result buffers, reused every step
    out_host = solver.alloc_outputs(B, traj, on_device=False)    # pinned host memory for the e2e leg

    def step_device():
        step_no[0] += 1
        raw = solver.solve_rows(rows_devs[step_no[0] % NBATCH], trajectories=traj, out=out_dev)
        if world > 1:   # the single allgather of per-problem results (tf, final mass, status, iterations)
            send = torch.stack([raw["tf"], raw["final_mass"], raw["status"].double(), raw["iterations"].double()], dim=1)
            dist.all_gather_into_tensor(gathered, send)
        return raw

    fp64_peak = solver.measure_fp64_peak()
    for _ in range(args.warmup):
        # identical to a timed step, including the host-side bookkeeping: the first use of each torch
        # reduction kernel costs ~0.3 s of lazy module loading that must not land in the timed region
        raw = step_device()
        solver.last_kernel_ms()
        int((raw["status"] == 0).sum()); int(raw["iterations"].sum())
    torch.cuda.synchronize()
    launches0 = solver.kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("LMATO_NO_CLOCKS") != "1":
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, conv, iters_total = 0.0, 0, 0
    ev0.record()
    for _ in range(args.steps):
        raw = step_device()
        kernel_ms += solver.last_kernel_ms()          # CUDA events around the IPM kernel on its stream
        conv += int((raw["status"] == 0).sum())
        iters_total += int(raw["iterations"].sum())
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.kernel_launches() - launches0
    stats = torch.tensor([ms, kernel_ms, float(conv), float(iters_total), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, kernel_ms = float(mx[0]), float(mx[1])
        conv_all, iters_all, launches_all = float(sm[2]), float(sm[3]), int(sm[4])
    else:
        conv_all, iters_all, launches_all = float(conv), float(iters_total), int(launches)
    value = conv_all / (ms * 1e-3)

    # ---- end-to-end through the host API (pinned host tensors in, pinned host tensors out) ----
    for j in range(2):
        solver.solve_rows(rows_hosts[j % NBATCH], trajectories=traj, out=out_host)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    conv_e = 0
    for j in range(args.steps):
        r = solver.solve_rows(rows_hosts[(j + 2) % NBATCH], trajectories=traj, out=out_host)   # synchronous: H2D + solve + D2H
        conv_e += int((r["status"] == 0).sum())
    e2e_s = time.perf_counter() - t0
    e = torch.tensor([e2e_s, float(conv_e)], dtype=torch.float64, device=dev)
    if world > 1:
        emx = e.clone(); dist.all_reduce(emx, op=dist.ReduceOp.MAX)
        esm = e.clone(); dist.all_reduce(esm, op=dist.ReduceOp.SUM)
        e2e_s, conv_e_all = float(emx[0]), float(esm[1])
    else:
        conv_e_all = float(conv_e)
    h2d = rows_host.numel() * 8
    d2h = (10 * nt * B * 8 if traj else 0) + B * (8 * 3 + 4 * 2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- latency of a single nominal solve through optimise() ----
    lat = []
    s1 = lm.AscentSolver(lm.Mesh(nt=nt), opts, device=dev)
    r1 = lm.AscentParams().rows(1).pin_memory()
    for i in range(103):        # 2 warm-up calls + 101 timed (median over >= 101 repeats)
        t0 = time.perf_counter(); s1.solve_rows(r1, trajectories=True); dt = time.perf_counter() - t0
        if i >= 2:
            lat.append(dt * 1e3)

    hbm_peak, peak_src = load_peaks()
    # Algorithmic work of ONE GPU's kernel launches (iterations are summed over ranks, kernel_ms is the
    # max over ranks): sum over its problems of iterations x stages x {336 B | 2147 FLOP}, per second of
    # kernel time (CUDA events on the launching stream).
    stages_gpu = (iters_all / world) * N
    ach_gbs_gpu = stages_gpu * BYTE_PER_STAGE / (kernel_ms * 1e-3) * 1e-9
    ach_gf_gpu = stages_gpu * FLOP_PER_STAGE / (kernel_ms * 1e-3) * 1e-9
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg4: elliptical ascent (Launch_Optimiser.py defaults), nt={nt}, NODES=2, "
                               f"6-parameter dispersions; {NBATCH} different draws (seed 11+rank+1000j) cycled over the steps",
                   "batch_per_gpu": B, "global_batch": B * world, "tol": opts.tol, "obj_scale": opts.obj_scale,
                   "dcost": 1e-5 if opts.dcost is None else opts.dcost, "warm_start": bool(opts.warm_start),
                   "trajectories": traj, "parallelism": (f"index-sharded x{world}, one allgather of results" + (f"; ranks bound to their GPU's NUMA node ({numa} CPUs)" if numa else "")) if world > 1 else "single GPU",
                   "l2": f"no flush needed: the kernel streams a {solver.workspace_bytes(B) / 2**30:.1f} GiB workspace (>> 126 MB L2) every sweep"},
        "converged_fraction": conv_all / (B * world * args.steps),
        "mean_iterations": iters_all / (B * world * args.steps),
        "kernel_ms_per_step": kernel_ms / args.steps,
        "p50_latency_ms": {"single_solve_host_api": statistics.median(lat), "batch_per_problem": ms / args.steps / B},
        "e2e": {"value": conv_e_all / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": ach_gbs_gpu, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach_gbs_gpu / hbm_peak,
                     "traffic": NCU_TRAFFIC_BYTES.get((bool((opts.dcost if opts.dcost is not None else 1e-5) > 0), B, nt)) if traj else None,
                     "traffic_source": "profiles/r01_ncu_ascent_ipm_kernel_metrics.csv (one launch, bytes)",
                     "peak_source": peak_src,
                     "kernel": "ascent_ipm_kernel", "algorithmic_bytes_per_stage_iter": BYTE_PER_STAGE,
                     "fp64": {"achieved_gflops": ach_gf_gpu, "peak_gflops": fp64_peak, "frac": ach_gf_gpu / fp64_peak,
                              "algorithmic_flop_per_stage_iter": FLOP_PER_STAGE,
                              "peak_source": "measured in this run (DFMA micro-benchmark, lmato_measure_fp64_peak)"}},
    }
    if world == 1 and not args.no_cpu_baseline:
        import numpy as np
        from oracle.cpu_baseline import OraclePool
        pool = OraclePool()
        n = min(2 * pool.cores, 256)
        _tf, st, it, wall = pool.solve(rows_host[:, :n].numpy(), nt, opts.tol, opts.obj_scale)
        pool.close()
        # parity spot-check of this very run: same inputs, GPU vs oracle
        gpu_tf = solver.solve_rows(rows_devs[0][:, :n].contiguous(), trajectories=False)["tf"].cpu().numpy()
        line["cpu_baseline"] = {"value": float((st == 0).sum() / wall), "unit": "solves/s", "cores": pool.cores, "kind": "port",
                                "sample": f"first {n} problems of this rank's batch, one problem per process on {pool.cores} cores",
                                "note": "oracle restatement (numpy/scipy sparse IPM), not GEKKO/IPOPT (absent from this image)",
                                "mean_iterations": float(np.mean(it)),
                                "max_rel_tf_diff_vs_gpu": float(np.max(np.abs(gpu_tf - _tf) / _tf))}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
