"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the pinned oracle).

CPU: the oracle must still reproduce them (drift guard: any change to oracle/ that moves a number shows up here).
GPU: the CUDA path, through the C ABI, is held to the same vectors at north_star's tolerances."""
import os

import numpy as np
import pytest

from limo_b200 import synth
from tests.golden import make_golden as mg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _check_solve(res, win, gold, pose_tol, cost_rel, iter_slack=0):
    solves = res.solves
    assert len(solves) == len(gold["termination"])
    assert [s.termination for s in solves] == list(gold["termination"])
    assert [s.num_landmarks for s in solves] == list(gold["num_landmarks"])
    assert [s.num_successful_steps for s in solves] == list(gold["num_successful_steps"])
    assert np.abs(np.array([s.num_iterations for s in solves]) - gold["num_iterations"]).max() <= iter_slack
    assert np.allclose([s.initial_cost for s in solves], gold["initial_cost"], rtol=cost_rel)
    assert np.allclose([s.final_cost for s in solves], gold["final_cost"], rtol=cost_rel, atol=1e-14)
    assert np.array_equal(res.lm_rejected[:win.n_lm], gold["lm_rejected"])
    assert np.linalg.norm(res.kf_pose[:, 4:] - gold["kf_pose"][:, 4:], axis=1).max() <= pose_tol
    assert np.abs(res.kf_pose[:, :4] - gold["kf_pose"][:, :4]).max() <= max(pose_tol, 1e-7)
    dl = np.linalg.norm(res.lm_pos[:win.n_lm] - gold["lm_pos"], axis=1)
    assert np.percentile(dl, 95) <= max(pose_tol, 1e-6) and dl.max() <= 0.1


@pytest.mark.parametrize("name", sorted(mg.SOLVE_CASES))
def test_oracle_matches_golden_solve(oracle, name):
    cfg, kw = mg.SOLVE_CASES[name]
    win = synth.make_window(cfg, **kw)
    _check_solve(oracle.solve_window(win, num_threads=1), win, _load("solve_%s.npz" % name), 1e-12, 1e-12)


@pytest.mark.parametrize("name", sorted(mg.EVAL_CASES))
def test_oracle_matches_golden_eval(oracle, name):
    cfg, kw = mg.EVAL_CASES[name]
    gold = _load("eval_%s.npz" % name)
    step = int(gold["step"][0])
    r, jp, jl, cost, failed = oracle.evaluate(synth.make_window(cfg, **kw))
    assert failed == gold["failed"][0]
    assert cost == pytest.approx(gold["cost"][0], rel=1e-13)
    assert np.allclose(r[::step], gold["r"], rtol=1e-13, atol=1e-13)
    assert np.allclose(jp[::step], gold["jp"], rtol=1e-12, atol=1e-12)
    assert np.allclose(jl[::step], gold["jl"], rtol=1e-12, atol=1e-12)


def test_oracle_matches_golden_lidar(oracle):
    cloud, T, K, feats = synth.make_lidar_scene()
    assert np.array_equal(oracle.lidar_depth(cloud, T, K, feats), _load("lidar_scene.npz")["depth"])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def handle():
    from limo_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.SOLVE_CASES))
def test_cuda_matches_golden_solve(handle, name):
    cfg, kw = mg.SOLVE_CASES[name]
    win = synth.make_window(cfg, **kw)
    # north_star: pose translation <= 1e-6 m, cost relative <= 1e-8 (FP64)
    _check_solve(handle.solve_window(win), win, _load("solve_%s.npz" % name), 1e-6, 1e-8, iter_slack=3 if cfg == 3 else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.EVAL_CASES))
def test_cuda_matches_golden_eval(handle, name):
    cfg, kw = mg.EVAL_CASES[name]
    gold = _load("eval_%s.npz" % name)
    step = int(gold["step"][0])
    r, jp, jl, cost, failed = handle.evaluate(synth.make_window(cfg, **kw))
    assert failed == gold["failed"][0]
    assert cost == pytest.approx(gold["cost"][0], rel=1e-12)
    assert np.allclose(r[::step], gold["r"], rtol=1e-11, atol=1e-11)
    assert np.allclose(jp[::step], gold["jp"], rtol=1e-10, atol=1e-9 * np.abs(gold["jp"]).max())
    assert np.allclose(jl[::step], gold["jl"], rtol=1e-10, atol=1e-9 * np.abs(gold["jl"]).max())


@pytest.mark.gpu
def test_cuda_matches_golden_lidar(handle):
    cloud, T, K, feats = synth.make_lidar_scene()
    d, _ = handle.lidar_depth(cloud, T, K, feats)
    assert np.array_equal(d, _load("lidar_scene.npz")["depth"])  # bit-exact (single precision, -fmad=false)
