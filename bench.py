#!/usr/bin/env python
"""bench.py -- BA windows/s of the B200 window solver on BASELINE.json's headline configuration.

One "step" = one complete trimmed bundle-adjustment solve (all LM iterations + trimming round, Ceres-equivalent
termination) of a BATCH of independent synthetic windows of config 2 (30 keyframes / 3000 landmarks / 40000
observations, mono + lidar depth, FP64).  `value` is whole-job windows/s with the batch resident in HBM; `e2e` is the
same metric through the host-buffer C-ABI path (pack + H2D + solve + D2H every step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
"""
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

import numpy as np  # noqa: E402

METRIC = "BA windows/s (30 KF, 3k LM, 40k obs)"
# Algorithmic bytes per observation of the residual/Jacobian kernel, mono + depth FP64: reads u, v, d (12) +
# keyframe index (4) + landmark data (~3); writes residual 3x8 + J_pose 3x6x8 [+ J_landmark 3x3x8].  Since round 2 the
# small-window path does not materialise J_landmark (its consumers re-form it as J_pose[:, 3:6] R): 259 - 72 = 187 B.
FUSED = os.environ.get("KBA_FUSED", "1") != "0"
B_OBS_ALGORITHMIC = 187.0 if FUSED else 259.0
# dram__bytes_read.sum + dram__bytes_write.sum of one k_eval_obs<true> launch / its observations, from the ncu --set full
# capture summarised in profiles/ (re-measured whenever the kernel changes)
B_OBS_DRAM_MEASURED = 277.0
TRAFFIC_SOURCE = "ncu --set full, profiles/r01_v11_ncu_summary.md: (0.200 GB read + 1.293 GB written) / 5.39 M observations"


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_windows(n_distinct, rank, config=2):
    from limo_b200 import parallel
    return parallel.windows_for_rank(n_distinct, rank, config)


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(windows, n_sample, threads):
    """the CPU oracle (port of the reference algorithm, oracle/) timed on the host cores"""
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    from oracle import oracle as orc
    orc.lib()
    orc.solve_window(windows[0], num_threads=threads)  # warm-up (page-in, thread pool)
    t = time.perf_counter()
    for i in range(n_sample):
        orc.solve_window(windows[i % len(windows)], num_threads=threads)
    dt = time.perf_counter() - t
    return n_sample / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (its restatement in oracle/, since Ceres is not installable
    offline -- see DESIGN.md) on the host cores, same metric and configuration."""
    if rank != 0:
        return
    threads = min(usable_cores(), 32)
    wins = make_windows(2, 0)
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    from oracle import oracle as orc
    orc.lib()
    per_step = 1
    for _ in range(max(args.warmup, 1)):
        orc.solve_window(wins[0], num_threads=threads)
    t = time.perf_counter()
    for s in range(args.steps):
        for i in range(per_step):
            orc.solve_window(wins[(s + i) % len(wins)], num_threads=threads)
    dt = time.perf_counter() - t
    val = args.steps * per_step / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config 2: 30 KF / 3000 LM / 40000 obs, mono + lidar depth, FP64; %d window per step"
                                  % per_step},
           "cpu_baseline": {"value": val, "unit": "windows/s", "cores": threads, "kind": "port",
                            "sample": "%d full window solves (oracle/, OpenMP %d threads)" % (args.steps * per_step, threads)},
           "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=296, help="windows per GPU per step (two per SM: the windows of a batch\n"
                    "advance in lock-step passes, a larger batch amortises the passes in which only the slowest windows are left)")
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic windows per GPU (tiled to --batch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=16, help="window solves timed for cpu_baseline (~0.7 s each)")
    ap.add_argument("--in-flight", type=int, default=4, help="steps in flight of the end-to-end measurement (handles / streams)")
    args = ap.parse_args()

    from limo_b200 import parallel
    rank, local_rank, world = parallel.rank_info()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the kba_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    parallel.init("nccl", torch.device("cuda", local_rank))
    from limo_b200 import capi
    # host threads that pack a step's windows: share the box's cores between the ranks and the steps in flight
    os.environ.setdefault("KBA_HOST_THREADS", str(max(2, min(16, usable_cores() // (max(1, min(args.in_flight, usable_cores() // (4 * world))) * world)))))

    n_distinct = max(1, min(args.distinct, args.batch))
    base = make_windows(n_distinct, rank)
    windows = [base[i % n_distinct] for i in range(args.batch)]
    n_obs_win = windows[0].n_obs

    stream = torch.cuda.current_stream()
    h = capi.Handle(local_rank, stream=stream.cuda_stream)
    opt = capi.default_options()
    batch = h.batch(windows)

    def barrier():
        parallel.barrier(cuda=True)

    # ---- warm-up ----
    for _ in range(args.warmup):
        batch.solve(opt)
    barrier()
    h.counters(reset=True)
    h.enable_kernel_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region: K steps, inputs resident in HBM ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        batch.solve(opt)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    cnt = h.counters(reset=True)
    h.enable_kernel_timing(False)
    results = batch.download()
    ok = all(r.c.status == 0 for r in results)
    iters = [sum(s.num_iterations for s in r.solves) for r in results]

    # ---- end-to-end: host buffers -> pack -> H2D -> solve -> D2H, every step ----
    for _ in range(1):
        batch.upload(); batch.solve(opt); batch.download(results=results)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record(stream)
    t_up = t_dn = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        batch.upload()       # host pack (threads) + async H2D
        t1 = time.perf_counter()
        batch.solve(opt)     # returns when every window is done
        t2 = time.perf_counter()
        batch.download(results=results)
        t_up += t1 - t0; t_dn += time.perf_counter() - t2
    e1.record(stream)
    barrier()
    ms_e2e_seq = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t_wall))

    # ---- end-to-end, several steps in flight: further handles (own stream, own device buffers) let the host pack and copy
    #      step i+1 while the GPU solves step i.  Every step still does its own pack + H2D + solve + D2H inside the
    #      timed region; ctypes releases the GIL during the C calls, so two Python threads are enough. ----
    # a lane needs a host thread for the launches plus a few packing threads: fewer lanes when the ranks share few cores
    n_lanes = max(1, min(args.in_flight, usable_cores() // (4 * world)))
    extra = []
    lanes = [(batch, results)]
    if n_lanes * world > 1:  # the additional handles sleep while they wait for the GPU (see kba_api.cu wait_stream)
        os.environ["KBA_BLOCKING_SYNC"] = "1"
    for _ in range(n_lanes - 1):
        st_ = torch.cuda.Stream()
        h_ = capi.Handle(local_rank, stream=st_.cuda_stream)
        b_ = h_.batch(windows)
        extra.append((st_, h_, b_))
        lanes.append((b_, b_.download()))
    for b_, r_ in lanes:  # warm-up of both lanes
        b_.upload(); b_.solve(opt); b_.download(results=r_)
    barrier()

    todo = {"left": args.steps}
    todo_lock = threading.Lock()

    def lane(idx):
        torch.cuda.set_device(local_rank)
        b_, r_ = lanes[idx]
        while True:
            with todo_lock:  # the lanes pull steps from one counter: balanced for any K
                if todo["left"] <= 0:
                    return
                todo["left"] -= 1
            b_.upload(); b_.solve(opt); b_.download(results=r_)

    threads = [threading.Thread(target=lane, args=(i,)) for i in range(min(n_lanes, args.steps))]
    t_wall = time.perf_counter()
    for t_ in threads:
        t_.start()
    for t_ in threads:
        t_.join()
    torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t_wall)
    barrier()
    ok = ok and all(r.c.status == 0 for _, rs in lanes[1:] for r in rs)
    clocks = sampler.stop()  # sampled over all timed regions
    h2d, d2h = batch.transfer_bytes()

    ms, ms_e2e, ms_e2e_seq = parallel.max_over_ranks([ms, ms_e2e, ms_e2e_seq], device="cuda")  # slowest rank

    if rank == 0:
        total_windows = world * args.batch * args.steps
        value = total_windows / (ms * 1e-3)
        peak, peak_src = measured_peak()
        jac_gbs = (cnt.jacobian_obs * B_OBS_ALGORITHMIC / (cnt.ms_jacobian * 1e-3) / 1e9) if cnt.ms_jacobian > 0 else None
        threads = min(usable_cores(), 32)
        cpu_val, cpu_dt = cpu_baseline(base, args.cpu_sample, threads)
        out = {
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: 30 KF / 3000 LM / 40000 obs, mono + lidar depth, FP64",
                       "batch_windows_per_gpu": args.batch, "distinct_windows_per_gpu": n_distinct,
                       "parallelism": "independent windows per GPU (no data-path collective)" if world > 1 else "1 GPU",
                       "lm_iterations_per_window_mean": float(np.mean(iters)),
                       "lm_iterations_per_window_max": int(np.max(iters)),
                       "l2_policy": "inputs larger than L2 (%.1f GB of Jacobian blocks per pass)"
                                    % (args.batch * n_obs_win * 240 / 1e9),
                       "all_windows_converged": bool(ok)},
            "e2e": {"value": total_windows / (ms_e2e * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps_in_flight": n_lanes, "sequential_value": total_windows / (ms_e2e_seq * 1e-3),
                    "host_pack_upload_ms_per_step": 1e3 * t_up / args.steps,
                    "download_ms_per_step": 1e3 * t_dn / args.steps},
            "gpu_launches": int(cnt.launches_total),
            "roofline": {"kernel": "k_eval_obs<true> (residual/Jacobian)", "bound": "hbm", "achieved": jac_gbs,
                         "peak": peak, "unit": "GB/s", "frac": (jac_gbs / peak) if jac_gbs else None,
                         "peak_source": peak_src,
                         "traffic": B_OBS_DRAM_MEASURED * cnt.jacobian_obs / max(cnt.launches_jacobian, 1),
                         "traffic_unit": "bytes per launch", "traffic_source": TRAFFIC_SOURCE,
                         "algorithmic_bytes_per_obs": B_OBS_ALGORITHMIC,
                         "launch_ms_mean": cnt.ms_jacobian / max(cnt.launches_jacobian, 1),
                         "launches": int(cnt.launches_jacobian), "share_of_timed_region": cnt.ms_jacobian / ms,
                         "obs_per_launch_mean": cnt.jacobian_obs / max(cnt.launches_jacobian, 1)},
            "host_cores": usable_cores(),
            "cpu_baseline": {"value": cpu_val, "unit": "windows/s", "cores": threads, "kind": "port",
                             "sample": "%d full window solves of the same workload, %.1f s (oracle/, OpenMP)"
                                       % (args.cpu_sample, cpu_dt)},
            "clocks": clocks,
        }
        print(json.dumps(out))
    for _, h_, b_ in extra:
        b_.close()
        h_.close()
    batch.close()
    h.close()
    parallel.finalize()


if __name__ == "__main__":
    main()
