"""worker of tests/test_parallel_gloo.py: one of WORLD_SIZE CPU processes (gloo) running the N > 1 plumbing of bench.py
with the oracle standing in for the GPU solver (tiny windows)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from limo_b200 import parallel, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, _, world = parallel.rank_info()
    assert parallel.init("gloo")
    orc.lib()
    # --- weak scaling: every rank owns its own windows (same shapes, different seeds) ---
    seeds = parallel.window_seeds(2, rank)
    wins = [synth.make_window(1, seed=s, n_kf=4, n_lm=60, n_obs=200) for s in seeds]
    parallel.barrier()
    t = time.perf_counter()
    costs = [orc.solve_window(w).solves[-1].final_cost for w in wins]
    ms = 1e3 * (time.perf_counter() - t) + 5.0 * rank  # rank-dependent so the max is distinguishable
    parallel.barrier()
    ms_max, = parallel.max_over_ranks([ms])
    n_windows, cost_sum = parallel.sum_over_ranks([len(wins), sum(costs)])
    all_ms = parallel.sum_over_ranks([ms if r == rank else 0.0 for r in range(world)])
    # --- one large window split by landmark blocks: the shards must tile it exactly ---
    big = synth.make_window(2, seed=99, n_kf=8, n_lm=240, n_obs=1600)
    sub, j0, j1 = parallel.shard_window(big, rank, world)
    n_obs_total, n_lm_total = parallel.sum_over_ranks([sub.n_obs, sub.n_lm])
    # the data every rank would contribute to the exchanged reduced system: here the per-keyframe observation counts
    counts = np.bincount(sub.obs_kf, minlength=big.n_kf).astype(float)
    counts_sum = parallel.sum_over_ranks(counts)
    # the packed scalar exchange of the sharded solve (kba_shard.cu / k_shard_scalars): ONE sum all-reduce carries the sums AND the
    # gradient max-norm, as one slot per rank (the others contribute zero) -- the maximum is taken locally afterwards
    g_mine = 0.25 + 0.5 * rank
    xs = np.zeros(16 + world)
    xs[:4] = [1.0 + rank, 2.0, 3.0 * rank, 4.0]
    xs[16 + rank] = g_mine
    xs_sum = np.asarray(parallel.sum_over_ranks(xs))
    packed_ok = bool(xs_sum[16:].max() == max(0.25 + 0.5 * r for r in range(world)) and xs_sum[0] == sum(1.0 + r for r in range(world))
                     and np.count_nonzero(xs_sum[4:16]) == 0)
    ok_slice = bool(np.array_equal(sub.lm_pos, big.lm_pos[j0:j1]) and np.array_equal(
        sub.obs_u, big.obs_u[big.lm_obs_ptr[j0]:big.lm_obs_ptr[j1]]))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(dict(world=world, ms_max=ms_max, all_ms=all_ms, n_windows=n_windows, cost_sum=cost_sum, seeds=seeds,
                           n_obs_total=n_obs_total, n_lm_total=n_lm_total, counts_sum=counts_sum,
                           counts_full=np.bincount(big.obs_kf, minlength=big.n_kf).tolist(),
                           big=(big.n_obs, big.n_lm), ok_slice=ok_slice, packed_ok=packed_ok, shard0=(j0, j1, sub.n_obs)), f)
    parallel.finalize()


if __name__ == "__main__":
    main()
