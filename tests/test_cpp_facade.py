"""The C++ drop-in facade (include/keyframe_bundle_adjustment/, limo_b200/csrc/facade/) against the reference's gtest
cases restated in tests/cpp/test_facade.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "limo_b200", "csrc"), "-s", "all", "facade"])


def test_facade_host_logic():
    """Triangulator.process, LandmarkCreator.CreateWithDepth, deactivateKeyframes, NotEnoughKeyframesException,
    cheirality selection, LandmarkSelector.voxel, the add-depth scheme -- no GPU needed"""
    _build()
    out = subprocess.run([EXE, "cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_facade_host_logic_strict_eigen():
    """the same cases on the strict build leg: the facade sources compiled with -DKBA_EIGEN_POISON_UNINIT, where the Eigen
    subset offers only spellings that exist in Eigen 3.3 and default-constructed vectors / matrices are NaN-filled (real
    Eigen leaves them uninitialised) -- an accumulation into a default-constructed object would poison the results"""
    _build()
    out = subprocess.run([EXE + "_strict", "cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_facade_solves_on_gpu():
    """KeyFrameBundleAdjustment.solve / solve_depth (mono + two-camera rigs, 3 noise settings each) and
    adjustMotionOnly through BundleAdjusterKeyframes::solve() / adjustPoseOnly() of the facade"""
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
