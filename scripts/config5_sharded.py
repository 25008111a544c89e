#!/usr/bin/env python
"""BASELINE config 5: ONE 100-keyframe window (20k landmarks, ~300k observations) solved on N GPUs, landmark blocks
sharded over the ranks, NCCL all-reduce of the reduced pose system per LM iteration (include/kba_b200.h, kba_shard.cu).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/config5_sharded.py [--steps K] [--check]

Prints one JSON line (rank 0): device ms per solve (max over ranks), and with --check the deviation from the
single-GPU solve of the whole window computed by rank 0 on its own GPU."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--small", action="store_true", help="40 keyframes / 3000 landmarks (quick test)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from limo_b200 import capi, parallel, synth
    rank, local_rank, world = parallel.rank_info()
    torch.cuda.set_device(local_rank)
    parallel.init("nccl", torch.device("cuda", local_rank))
    win = synth.make_window(5, n_kf=40, n_lm=3000, n_obs=45000) if args.small else synth.make_window(5)
    sub, j0, j1 = parallel.shard_window(win, rank, world)
    torch.cuda.set_stream(torch.cuda.Stream())  # not the legacy stream: the library captures the passes into a CUDA graph
    stream = torch.cuda.current_stream()
    h = capi.Handle(local_rank, stream=stream.cuda_stream)
    # NCCL id: rank 0 creates, torch.distributed broadcasts the 128 bytes
    idt = torch.zeros(capi.SHARD_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.shard_unique_id()), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(idt, 0)
    comm = capi.ShardComm(h, rank, world, bytes(idt.cpu().numpy().tobytes()))
    batch = h.batch([sub])
    batch.set_shard(comm, j0, win.n_lm)
    opt = capi.default_options()
    res = None
    times = []
    for step in range(args.steps + 1):  # first one is the warm-up
        batch.upload()
        parallel.barrier(cuda=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.solve(opt)
        e1.record(stream)
        parallel.barrier(cuda=True)
        ms, = parallel.max_over_ranks([e0.elapsed_time(e1)], device="cuda")
        if step > 0:
            times.append(ms)
        res = batch.download(results=res)
    r = res[0]
    out = {"workload": "config 5 sharded: %d KF / %d LM / %d obs" % (win.n_kf, win.n_lm, win.n_obs), "n_gpus": world,
           "ms_per_solve": float(np.median(times)), "iterations": [s.num_iterations for s in r.solves],
           "final_cost": r.solves[-1].final_cost, "landmarks_of_rank0": int(sub.n_lm)}
    # gather the landmark blocks and the rejection flags on rank 0
    n_rej = parallel.sum_over_ranks([int(r.lm_rejected[:sub.n_lm].sum())], device="cuda")[0]
    out["rejected"] = int(n_rej)
    if args.check:
        lm_full = torch.zeros(win.n_lm * 3, dtype=torch.float64, device="cuda")
        lm_full[3 * j0:3 * j1] = torch.from_numpy(r.lm_pos[:sub.n_lm].reshape(-1)).cuda()
        if world > 1:
            dist.all_reduce(lm_full)
        if rank == 0:
            h1 = capi.Handle(local_rank, stream=stream.cuda_stream)
            ref = h1.solve_window(win)
            out["single_gpu_ms"] = 1e3 * ref.c.time_sec
            out["single_gpu_iterations"] = [s.num_iterations for s in ref.solves]
            out["max_dt_vs_single_gpu"] = float(np.linalg.norm(r.kf_pose[:, 4:] - ref.kf_pose[:, 4:], axis=1).max())
            out["rel_dcost_vs_single_gpu"] = float(abs(r.solves[-1].final_cost - ref.solves[-1].final_cost) / ref.solves[-1].final_cost)
            dl = np.linalg.norm(lm_full.cpu().numpy().reshape(-1, 3) - ref.lm_pos[:win.n_lm], axis=1)
            out["p95_dlm_vs_single_gpu"] = float(np.percentile(dl, 95))
            out["rejected_single_gpu"] = int(ref.lm_rejected[:win.n_lm].sum())
            h1.close()
    if rank == 0:
        print(json.dumps(out))
    batch.close()
    comm.close()
    h.close()
    parallel.finalize()


if __name__ == "__main__":
    main()
