"""ctypes binding of the CPU oracle (oracle/libkba_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (limo_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from limo_b200.capi_types import (KbaEvalOut, KbaLidarOptions, KbaOptions, KbaResult, KbaWindow, Result, c_double_p, c_int32_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkba_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "kba_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def use_native_build():
    """bench.py's CPU legs: compile the oracle with -O3 -march=native for the host this process runs
    on -- into a temporary directory, so that a library tuned for one machine never travels to another -- and load that
    instead of the portable -O2 build.  Must be called before the first lib().  Falls back to the portable build (and
    says so) when the compiler is missing.  Returns the description of the build for the bench line."""
    global _LIB_PATH
    import tempfile
    if _lib is not None:
        raise RuntimeError("use_native_build() must be called before the oracle library is loaded")
    out = os.path.join(tempfile.mkdtemp(prefix="kba_oracle_native_"), "libkba_oracle_native.so")
    # -ffp-contract=off: same arithmetic as the portable build (the lidar oracle is compared bit for bit)
    cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-std=gnu11", "-fPIC", "-fopenmp", "-I" + os.path.join(_HERE, "..", "include"), "-shared",
           "-o", out, os.path.join(_HERE, "kba_oracle.c"), os.path.join(_HERE, "lidar_oracle.c"), "-lm"]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        return "gcc -O2 (portable build; native build failed)"
    _LIB_PATH = out
    return "gcc -O3 -march=native -ffp-contract=off -fopenmp, built on this host"


def lib():
    global _lib
    if _lib is None:
        if _LIB_PATH.endswith("libkba_oracle.so"):
            build()
        _lib = C.CDLL(_LIB_PATH)
        dp = c_double_p
        _lib.kbo_solve_window.argtypes = [C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaResult), C.c_int]
        _lib.kbo_solve_window.restype = C.c_int
        _lib.kbo_eval.argtypes = [C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaEvalOut)]
        _lib.kbo_eval.restype = C.c_int
        _lib.kbo_default_options.argtypes = [C.POINTER(KbaOptions)]
        _lib.kbo_reprojection.argtypes = [dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp]
        _lib.kbo_reprojection.restype = C.c_int
        _lib.kbo_depth.argtypes = [dp, dp, dp, C.c_double, dp, dp, dp]
        _lib.kbo_gp_height.argtypes = [dp, dp, C.c_double, dp, dp, dp, dp, dp, dp]
        _lib.kbo_gp_motion.argtypes = [dp, dp, dp, dp, dp, dp, dp]
        _lib.kbo_scale_reg.argtypes = [dp, dp, C.c_double, dp, dp, dp]
        _lib.kbo_speed_reg.argtypes = [dp, dp, C.c_double, dp, dp, dp]
        _lib.kbo_pose_plus.argtypes = [dp, dp, dp]
        _lib.kbo_dir_plus.argtypes = [dp, dp, dp]
        _lib.kbo_trimmer_quantile.argtypes = [dp, C.c_int, C.c_double, C.POINTER(C.c_uint8)]
        _lib.kbo_trimmer_quantile.restype = C.c_int
        _lib.kbo_trimmer_fix.argtypes = [dp, C.c_int, C.c_double, C.POINTER(C.c_uint8)]
        _lib.kbo_trimmer_fix.restype = C.c_int
        _lib.kbo_triangulate_rays.argtypes = [C.c_int, dp, dp, dp, dp]
        fp = C.POINTER(C.c_float)
        _lib.kbo_lidar_default_options.argtypes = [C.POINTER(KbaLidarOptions)]
        _lib.kbo_lidar_depth.argtypes = [fp, C.c_int, C.c_int, dp, dp, fp, C.c_int, C.POINTER(KbaLidarOptions), fp]
        _lib.kbo_lidar_depth.restype = C.c_int
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_double_p)


def default_options():
    o = KbaOptions()
    lib().kbo_default_options(C.byref(o))
    return o


def solve_window(win, opt=None, num_threads=1, iterations_capacity=256):
    opt = opt or default_options()
    res = Result(win, iterations_capacity)
    rc = lib().kbo_solve_window(C.byref(win.c), C.byref(opt), C.byref(res.c), num_threads)
    if rc != 0:
        raise RuntimeError("kbo_solve_window failed: %d" % rc)
    return res


def evaluate(win, opt=None):
    opt = opt or default_options()
    n = max(win.n_obs, 1)
    r = np.zeros((n, 3)); jp = np.zeros((n, 3, 6)); jl = np.zeros((n, 3, 3)); cost = np.zeros(1)
    failed = np.zeros(1, dtype=np.int32)
    out = KbaEvalOut()
    out.residual = r.ctypes.data_as(c_double_p)
    out.jac_pose = jp.ctypes.data_as(c_double_p)
    out.jac_lm = jl.ctypes.data_as(c_double_p)
    out.cost = cost.ctypes.data_as(c_double_p)
    out.failed = failed.ctypes.data_as(c_int32_p)
    rc = lib().kbo_eval(C.byref(win.c), C.byref(opt), C.byref(out))
    if rc != 0:
        raise RuntimeError("kbo_eval failed: %d" % rc)
    return r[:win.n_obs], jp[:win.n_obs], jl[:win.n_obs], float(cost[0]), int(failed[0])


# ---- single residual functions -------------------------------------------------------------------
def reprojection(pose, cam_pose, intr, point, u, v, jac=True):
    _, pp = _d(pose); _, cp = _d(cam_pose); _, ip = _d(intr); _, xp = _d(point)
    res = np.zeros(2); jp = np.zeros((2, 6)); jl = np.zeros((2, 3))
    ok = lib().kbo_reprojection(pp, cp, ip, xp, float(u), float(v), res.ctypes.data_as(c_double_p),
                                jp.ctypes.data_as(c_double_p) if jac else None,
                                jl.ctypes.data_as(c_double_p) if jac else None)
    return bool(ok), res, jp, jl


def depth(pose, cam_pose, point, d):
    _, pp = _d(pose); _, cp = _d(cam_pose); _, xp = _d(point)
    res = np.zeros(1); jp = np.zeros((1, 6)); jl = np.zeros((1, 3))
    lib().kbo_depth(pp, cp, xp, float(d), res.ctypes.data_as(c_double_p), jp.ctypes.data_as(c_double_p),
                    jl.ctypes.data_as(c_double_p))
    return res, jp, jl


def gp_height(pose, direction, dist, point):
    _, pp = _d(pose); _, dp_ = _d(direction); _, xp = _d(point)
    res = np.zeros(1); jp = np.zeros((1, 6)); jd = np.zeros((1, 3)); jdist = np.zeros((1, 1)); jl = np.zeros((1, 3))
    lib().kbo_gp_height(pp, dp_, float(dist), xp, res.ctypes.data_as(c_double_p), jp.ctypes.data_as(c_double_p),
                        jd.ctypes.data_as(c_double_p), jdist.ctypes.data_as(c_double_p), jl.ctypes.data_as(c_double_p))
    return res, jp, jd, jdist, jl


def gp_motion(pose0, pose1, dir0):
    _, p0 = _d(pose0); _, p1 = _d(pose1); _, d0 = _d(dir0)
    res = np.zeros(1); j0 = np.zeros((1, 6)); j1 = np.zeros((1, 6)); jd = np.zeros((1, 3))
    lib().kbo_gp_motion(p0, p1, d0, res.ctypes.data_as(c_double_p), j0.ctypes.data_as(c_double_p),
                        j1.ctypes.data_as(c_double_p), jd.ctypes.data_as(c_double_p))
    return res, j0, j1, jd


def scale_reg(pose1, pose0, scale):
    _, p1 = _d(pose1); _, p0 = _d(pose0)
    res = np.zeros(1); j1 = np.zeros((1, 6)); j0 = np.zeros((1, 6))
    lib().kbo_scale_reg(p1, p0, float(scale), res.ctypes.data_as(c_double_p), j1.ctypes.data_as(c_double_p),
                        j0.ctypes.data_as(c_double_p))
    return res, j1, j0


def speed_reg(pose, T_origin_before, dt, v_before):
    _, pp = _d(pose); _, tp = _d(T_origin_before); _, vp = _d(v_before)
    res = np.zeros(3); jp = np.zeros((3, 6))
    lib().kbo_speed_reg(pp, tp, float(dt), vp, res.ctypes.data_as(c_double_p), jp.ctypes.data_as(c_double_p))
    return res, jp


def pose_plus(pose, delta):
    _, pp = _d(pose); _, dp_ = _d(delta)
    out = np.zeros(7)
    lib().kbo_pose_plus(pp, dp_, out.ctypes.data_as(c_double_p))
    return out


def dir_plus(n, delta):
    _, np_ = _d(n); _, dp_ = _d(delta)
    out = np.zeros(3)
    lib().kbo_dir_plus(np_, dp_, out.ctypes.data_as(c_double_p))
    return out


def trimmer_quantile(values, q):
    v, vp = _d(values)
    rej = np.zeros(max(len(v), 1), dtype=np.uint8)
    n = lib().kbo_trimmer_quantile(vp, len(v), float(q), rej.ctypes.data_as(C.POINTER(C.c_uint8)))
    return n, rej[:len(v)].astype(bool)


def trimmer_fix(values, threshold):
    v, vp = _d(values)
    rej = np.zeros(max(len(v), 1), dtype=np.uint8)
    n = lib().kbo_trimmer_fix(vp, len(v), float(threshold), rej.ctypes.data_as(C.POINTER(C.c_uint8)))
    return n, rej[:len(v)].astype(bool)


def triangulate_rays(R_oc, t_oc, rays):
    R, Rp = _d(R_oc); t, tp = _d(t_oc); r, rp = _d(rays)
    out = np.zeros(3)
    lib().kbo_triangulate_rays(len(t.reshape(-1, 3)), Rp, tp, rp, out.ctypes.data_as(c_double_p))
    return out


def init_landmarks(win):
    """push() landmark initialisation for every landmark of `win`: (positions [n_lm, 3], flags [n_lm] uint8)"""
    pos = np.zeros((max(win.n_lm, 1), 3))
    flags = np.zeros(max(win.n_lm, 1), dtype=np.uint8)
    L = lib()
    L.kbo_init_landmarks.argtypes = [C.POINTER(type(win.c)), c_double_p, C.POINTER(C.c_uint8)]
    L.kbo_init_landmarks.restype = None
    L.kbo_init_landmarks(C.byref(win.c), pos.ctypes.data_as(c_double_p), flags.ctypes.data_as(C.POINTER(C.c_uint8)))
    return pos[:win.n_lm], flags[:win.n_lm]


def lidar_default_options():
    o = KbaLidarOptions()
    lib().kbo_lidar_default_options(C.byref(o))
    return o


def lidar_depth(cloud, T_cam_lidar, intr, features_uv, opt=None):
    cloud = np.ascontiguousarray(cloud, dtype=np.float32)
    feats = np.ascontiguousarray(features_uv, dtype=np.float32).reshape(-1, 2)
    T, Tp = _d(T_cam_lidar); K, Kp = _d(intr)
    out = np.zeros(max(len(feats), 1), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    rc = lib().kbo_lidar_depth(cloud.ctypes.data_as(fp), cloud.shape[0], cloud.shape[1], Tp, Kp, feats.ctypes.data_as(fp),
                               len(feats), C.byref(opt or lidar_default_options()), out.ctypes.data_as(fp))
    if rc != 0:
        raise RuntimeError("kbo_lidar_depth failed: %d" % rc)
    return out[:len(feats)]


class OracleBackend:
    """Drop-in for limo_b200.capi.Handle in host-logic tests: runs the CPU oracle instead of the CUDA library."""

    def __init__(self, num_threads=1):
        self.num_threads = num_threads

    def default_options(self):
        return default_options()

    def solve_window(self, win, opt=None):
        return solve_window(win, opt, self.num_threads)
