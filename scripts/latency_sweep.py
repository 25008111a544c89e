#!/usr/bin/env python
"""Latency / throughput of config-2 batches under the current KBA_GRAPH mode (read once per process by the library):
   python scripts/latency_sweep.py [batch ...]   -> one JSON line per batch (resident ms per solve, end-to-end ms, launches)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from limo_b200 import capi, parallel  # noqa: E402

batches = [int(a) for a in sys.argv[1:]] or [1, 64]
base = parallel.windows_for_rank(16, 0, 2)
torch.cuda.set_stream(torch.cuda.Stream())
stream = torch.cuda.current_stream()
h = capi.Handle(0, stream=stream.cuda_stream)
opt = capi.default_options()
for b in batches:
    wins = [base[i % len(base)] for i in range(b)]
    batch = h.batch(wins)
    steps = 20 if b == 1 else 5
    for _ in range(3):
        batch.solve(opt)
    h.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        batch.solve(opt)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t) / steps
    ms = e0.elapsed_time(e1) / steps
    cnt = h.counters(reset=True)
    res = None
    t = time.perf_counter()
    for _ in range(steps):
        batch.upload(); batch.solve(opt); res = batch.download(results=res)
    e2e = 1e3 * (time.perf_counter() - t) / steps
    print(json.dumps({"graph_mode": os.environ.get("KBA_GRAPH", "default"), "batch": b, "ms_per_solve": round(ms, 3),
                      "wall_ms_per_solve": round(wall, 3), "e2e_ms": round(e2e, 3), "windows_per_s": round(b / (ms * 1e-3), 1),
                      "launches_per_solve": cnt.launches_total / steps, "done": all(r.c.status == 0 for r in res),
                      "iters": [sum(s.num_iterations for s in r.solves) for r in res[:2]],
                      "final_cost0": res[0].c.final_cost}))
    batch.close()
h.close()
