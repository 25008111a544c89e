"""single-GPU solve of the 100-keyframe window (BASELINE config 5) -- used under ncu for the launch list"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limo_b200 import synth, capi
win = synth.make_window(5)
h = capi.Handle(0)
for _ in range(int(os.environ.get("REPS", "1"))):
    t = time.time(); r = h.solve_window(win); print("solve %.4f s (device %.4f s)" % (time.time() - t, r.c.time_sec), [s.num_iterations for s in r.solves])
h.close()
