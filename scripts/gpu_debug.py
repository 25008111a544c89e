"""Side-by-side iteration log: CUDA path vs CPU oracle (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from limo_b200 import synth, capi
from oracle import oracle as orc

def show(win, label):
    h = capi.Handle(0)
    t = time.time(); rg = h.solve_window(win); tg = time.time() - t
    t = time.time(); rc = orc.solve_window(win, num_threads=int(os.environ.get('KBA_ORACLE_THREADS', '8'))); tc = time.time() - t
    print("==== %s: n_kf %d n_lm %d n_obs %d | gpu %.4fs (device %.4fs) cpu %.3fs" % (label, win.n_kf, win.n_lm, win.n_obs, tg, rg.c.time_sec, tc))
    print(" gpu solves", [(s.initial_cost, s.final_cost, s.num_iterations, s.num_successful_steps, s.termination, s.num_landmarks, s.num_residual_blocks) for s in rg.solves], "status", rg.c.status)
    print(" cpu solves", [(s.initial_cost, s.final_cost, s.num_iterations, s.num_successful_steps, s.termination, s.num_landmarks, s.num_residual_blocks) for s in rc.solves])
    ig, ic = rg.iterations, rc.iterations
    for i in range(max(len(ig), len(ic))):
        a = ig[i] if i < len(ig) else None
        b = ic[i] if i < len(ic) else None
        f = lambda e: "s%d it%2d cost %.10e dc %+.3e g %.3e st %.3e rho %+.4f rad %.2e v%d ok%d" % (e.solve_index, e.iteration, e.cost, e.cost_change, e.gradient_max_norm, e.step_norm, e.relative_decrease, e.trust_region_radius, e.step_is_valid, e.step_is_successful) if e else "-"
        print("  G", f(a)); print("  C", f(b))
        if i > int(os.environ.get('KBA_DBG_ROWS', '14')): break
    print(" max |dt| %.3e  max |dq| %.3e  max |dlm| %.3e  rejected gpu %d cpu %d same %s" % (
        np.linalg.norm(rg.kf_pose[:, 4:] - rc.kf_pose[:, 4:], axis=1).max(), np.abs(rg.kf_pose[:, :4] - rc.kf_pose[:, :4]).max(),
        np.linalg.norm(rg.lm_pos[:win.n_lm] - rc.lm_pos[:win.n_lm], axis=1).max(), rg.lm_rejected.sum(), rc.lm_rejected.sum(),
        np.array_equal(rg.lm_rejected, rc.lm_rejected)))
    c = h.counters(); print(" launches", c.launches_total)
    h.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2"]
    if "1" in which: show(synth.make_window(1), "config1")
    if "s" in which: show(synth.make_window(2, n_kf=10, n_lm=300, n_obs=2500, seed=5), "small2")
    if "2" in which: show(synth.make_window(2), "config2")
    if "3" in which: show(synth.make_window(3, seed=41), "config3")
    if "3s" in which: show(synth.make_window(3, seed=41, n_kf=14, n_lm=500, n_obs=4500), "config3-small")
    if "5" in which: show(synth.make_window(5), "config5")
    if "5s" in which: show(synth.make_window(5, n_kf=40, n_lm=3000, n_obs=45000), "config5-small")
