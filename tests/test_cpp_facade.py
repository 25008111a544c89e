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
    cheirality selection -- no GPU needed"""
    _build()
    out = subprocess.run([EXE, "cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_facade_solves_on_gpu():
    """KeyFrameBundleAdjustment.solve / solve_depth (mono + two-camera rigs, 3 noise settings each) and
    adjustMotionOnly through BundleAdjusterKeyframes::solve() / adjustPoseOnly() of the facade"""
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
