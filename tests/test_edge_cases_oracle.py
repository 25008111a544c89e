"""The edge-case windows of tests/edge_windows.py through the CPU oracle: they pin the BEHAVIOUR the CUDA path is then
held to on the GPU box (tests/test_gpu_parity.py::test_edge_case_windows_match_oracle)."""
import numpy as np

from tests import edge_windows as ew

FAILURE, CONVERGENCE, NO_CONVERGENCE = 2, 0, 1


def test_ragged_rows(oracle):
    win = ew.ragged()
    r = oracle.solve_window(win)
    empty = np.diff(win.lm_obs_ptr) == 0
    assert empty.sum() == 4
    assert np.array_equal(r.lm_pos[:win.n_lm][empty], win.lm_pos[empty])   # not in the program: untouched
    assert r.solves[0].num_landmarks == win.n_lm - 4
    assert r.c.final_cost < r.c.initial_cost


def test_all_keyframes_fixed(oracle):
    win = ew.all_keyframes_fixed()
    r = oracle.solve_window(win)
    assert np.array_equal(r.kf_pose, win.kf_pose)
    assert r.c.final_cost < r.c.initial_cost and not np.array_equal(r.lm_pos[:win.n_lm], win.lm_pos)


def test_evaluation_failure_then_trimming(oracle):
    win = ew.evaluation_failure()
    r = oracle.solve_window(win)
    assert r.solves[0].termination == FAILURE and r.solves[0].num_iterations == 0
    assert r.lm_rejected[0] == 1                      # the landmark in the |z| < 0.01 band goes with the first trim
    assert r.solves[-1].termination == CONVERGENCE and r.solves[-1].final_cost < r.solves[-1].initial_cost


def test_too_small_for_trimming(oracle):
    win = ew.tiny()
    r = oracle.solve_window(win)
    assert r.c.num_solves == 1 and r.lm_rejected[:win.n_lm].sum() == 0
