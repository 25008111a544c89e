"""ctypes binding of the CUDA library limo_b200/libkba_b200.so (C ABI: include/kba_b200.h).

The product path: there is no CPU fallback.  Importing works without a GPU (symbols can be inspected), but every
computing call fails with KBA_ERR_CUDA when no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

from .capi_types import (KbaCounters, KbaEvalOut, KbaLidarOptions, KbaOptions, KbaResult, KbaTrackCaps, KbaWindow, Result, Window,
                         c_double_p, c_int32_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KBA_LIB_PATH") or os.path.join(_HERE, "libkba_b200.so")  # KBA_LIB_PATH: instrumented builds
_lib = None

SYMBOLS = ["kba_version", "kba_last_error", "kba_default_options", "kba_create", "kba_destroy", "kba_set_stream",
           "kba_solve_window", "kba_solve_batch", "kba_eval", "kba_batch_create", "kba_batch_upload",
           "kba_batch_solve", "kba_batch_download", "kba_batch_transfer_bytes", "kba_batch_jacobian_pass", "kba_batch_destroy",
           "kba_get_counters", "kba_enable_kernel_timing", "kba_lidar_default_options", "kba_lidar_depth",
           "kba_shard_unique_id", "kba_shard_comm_create", "kba_shard_comm_destroy", "kba_batch_set_shard",
           "kba_init_landmarks", "kba_track_create", "kba_track_destroy", "kba_track_push_keyframe", "kba_track_drop_keyframe",
           "kba_track_set_landmarks", "kba_track_set_keyframe_pose", "kba_track_set_keyframe_poses", "kba_track_solve", "kba_track_transfer_bytes"]


class KbaError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KbaError("CUDA library %s is missing: build it with `make -C limo_b200/csrc` (there is no CPU "
                           "fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.kba_version.restype = C.c_int
        L.kba_last_error.restype = C.c_char_p
        L.kba_default_options.argtypes = [C.POINTER(KbaOptions)]
        L.kba_create.argtypes = [C.POINTER(vp), C.c_int]
        L.kba_destroy.argtypes = [vp]
        L.kba_destroy.restype = None
        L.kba_set_stream.argtypes = [vp, vp]
        L.kba_solve_window.argtypes = [vp, C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaResult)]
        L.kba_solve_batch.argtypes = [vp, C.c_int32, C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaResult)]
        L.kba_eval.argtypes = [vp, C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaEvalOut)]
        L.kba_batch_create.argtypes = [vp, C.c_int32, C.POINTER(KbaWindow), C.POINTER(vp)]
        L.kba_batch_upload.argtypes = [vp, C.c_int32, C.POINTER(KbaWindow)]
        L.kba_batch_solve.argtypes = [vp, C.POINTER(KbaOptions)]
        L.kba_batch_download.argtypes = [vp, C.POINTER(KbaResult)]
        L.kba_batch_jacobian_pass.argtypes = [vp, C.POINTER(KbaOptions), C.c_int32, C.POINTER(C.c_float)]
        L.kba_batch_transfer_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.kba_batch_destroy.argtypes = [vp]
        L.kba_batch_destroy.restype = None
        L.kba_get_counters.argtypes = [vp, C.POINTER(KbaCounters), C.c_int]
        L.kba_enable_kernel_timing.argtypes = [vp, C.c_int]
        L.kba_shard_unique_id.argtypes = [C.c_char_p]
        L.kba_shard_comm_create.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(vp)]
        L.kba_shard_comm_destroy.argtypes = [vp]
        L.kba_shard_comm_destroy.restype = None
        L.kba_batch_set_shard.argtypes = [vp, vp, C.c_int32, C.c_int32]
        L.kba_init_landmarks.argtypes = [vp, C.POINTER(KbaWindow), c_double_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
        ip, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.kba_track_create.argtypes = [vp, C.POINTER(KbaTrackCaps), C.c_int32, c_double_p, c_double_p, C.POINTER(vp)]
        L.kba_track_destroy.argtypes = [vp]
        L.kba_track_destroy.restype = None
        L.kba_track_push_keyframe.argtypes = [vp, C.c_int32, c_double_p, c_double_p, C.c_int32, ip, ip, C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.kba_track_drop_keyframe.argtypes = [vp, C.c_int32]
        L.kba_track_set_landmarks.argtypes = [vp, C.c_int32, ip, c_double_p, c_double_p]
        L.kba_track_set_keyframe_pose.argtypes = [vp, C.c_int32, c_double_p, c_double_p]
        L.kba_track_set_keyframe_poses.argtypes = [vp, C.c_int32, ip, c_double_p, c_double_p]
        L.kba_track_solve.argtypes = [vp, C.c_int32, ip, u8p, C.c_int32, ip, C.POINTER(KbaWindow), C.POINTER(KbaOptions), C.POINTER(KbaResult)]
        L.kba_track_transfer_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.kba_lidar_default_options.argtypes = [C.POINTER(KbaLidarOptions)]
        L.kba_lidar_default_options.restype = None
        fp = C.POINTER(C.c_float)
        L.kba_lidar_depth.argtypes = [vp, fp, C.c_int32, C.c_int32, c_double_p, c_double_p, fp, C.c_int32,
                                      C.POINTER(KbaLidarOptions), fp, fp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise KbaError("kba_b200 error %d: %s" % (rc, lib().kba_last_error().decode()))


def default_options():
    o = KbaOptions()
    lib().kba_default_options(C.byref(o))
    return o


def lidar_default_options():
    o = KbaLidarOptions()
    lib().kba_lidar_default_options(C.byref(o))
    return o


class Batch:
    """Windows resident in HBM (kba_batch_*)."""

    def __init__(self, handle, windows):
        self.handle, self.windows = handle, list(windows)
        self._arr = (KbaWindow * len(self.windows))(*[w.c for w in self.windows])
        self._p = C.c_void_p()
        _check(lib().kba_batch_create(handle._p, len(self.windows), self._arr, C.byref(self._p)))

    def upload(self):
        _check(lib().kba_batch_upload(self._p, len(self.windows), self._arr))

    def solve(self, opt=None):
        _check(lib().kba_batch_solve(self._p, C.byref(opt or default_options())))

    def download(self, iterations_capacity=0, results=None):
        """results: reuse the buffers of an earlier download (avoids re-allocating numpy arrays every step)"""
        if results is None:
            results = [Result(w, max(iterations_capacity, 1)) for w in self.windows]
            self._res_arr = (KbaResult * len(results))(*[r.c for r in results])
        arr = self._res_arr
        _check(lib().kba_batch_download(self._p, arr))
        for r, c in zip(results, arr):
            r.c = c
        return results

    def set_shard(self, comm, lm_begin, lm_total):
        """this batch holds one rank's shard of a window split by landmark blocks; solve() becomes a collective call"""
        _check(lib().kba_batch_set_shard(self._p, comm._p, int(lm_begin), int(lm_total)))
        self._comm = comm

    def transfer_bytes(self):
        """(host->device bytes of the last upload, device->host bytes of the last download)"""
        a, b = C.c_int64(), C.c_int64()
        _check(lib().kba_batch_transfer_bytes(self._p, C.byref(a), C.byref(b)))
        return a.value, b.value

    def jacobian_pass(self, opt=None, repeats=1):
        ms = C.c_float()
        _check(lib().kba_batch_jacobian_pass(self._p, C.byref(opt or default_options()), repeats, C.byref(ms)))
        return ms.value

    def close(self):
        if self._p:
            lib().kba_batch_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Track:
    """Persistent, device-resident sliding window (kba_track_*): keyframes are uploaded once when pushed, a solve sends only
    the lists of active keyframe slots and selected landmark slots."""

    def __init__(self, handle, cam_intr, cam_pose, max_keyframes, max_landmarks, max_measurements, win_keyframes,
                 win_landmarks, win_observations, win_ground=0):
        self.handle = handle
        caps = KbaTrackCaps(max_keyframes, max_landmarks, max_measurements, win_keyframes, win_landmarks, win_observations, win_ground)
        intr = np.ascontiguousarray(cam_intr, dtype=np.float64).reshape(-1, 3)
        pose = np.ascontiguousarray(cam_pose, dtype=np.float64).reshape(-1, 7)
        self._p = C.c_void_p()
        _check(lib().kba_track_create(handle._p, C.byref(caps), len(intr), intr.ctypes.data_as(c_double_p),
                                      pose.ctypes.data_as(c_double_p), C.byref(self._p)))

    @staticmethod
    def _i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a, a.ctypes.data_as(C.POINTER(C.c_int32))

    def push_keyframe(self, slot, pose7, lm_slot, u, v, d, cam=None, plane4=None):
        fp = C.POINTER(C.c_float)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        pl = None if plane4 is None else np.ascontiguousarray(plane4, dtype=np.float64)
        lm, lmp = self._i32(lm_slot)
        cm, cmp_ = (None, C.cast(None, C.POINTER(C.c_int32))) if cam is None else self._i32(cam)
        uu, vv, dd = (np.ascontiguousarray(x, dtype=np.float32) for x in (u, v, d))
        _check(lib().kba_track_push_keyframe(self._p, int(slot), pose.ctypes.data_as(c_double_p),
                                             C.cast(None, c_double_p) if pl is None else pl.ctypes.data_as(c_double_p), len(lm), lmp, cmp_,
                                             uu.ctypes.data_as(fp), vv.ctypes.data_as(fp), dd.ctypes.data_as(fp)))

    def drop_keyframe(self, slot):
        _check(lib().kba_track_drop_keyframe(self._p, int(slot)))

    def set_landmarks(self, lm_slot, pos=None, weight=None):
        lm, lmp = self._i32(lm_slot)
        p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        _check(lib().kba_track_set_landmarks(self._p, len(lm), lmp, C.cast(None, c_double_p) if p is None else p.ctypes.data_as(c_double_p),
                                             C.cast(None, c_double_p) if w is None else w.ctypes.data_as(c_double_p)))

    def solve(self, kf_slots, kf_fixed, lm_slots, opt=None, **scalars):
        """scalars: scale_kf0, scale_kf1, scale_weight, scale_value, plane_reg_weight, plane_dist_fixed, gp_lm, gp_kf, gp_weight"""
        kf, kfp = self._i32(kf_slots)
        lm, lmp = self._i32(lm_slots)
        fx = np.ascontiguousarray(kf_fixed, dtype=np.uint8)
        n_kf, n_lm = len(kf), len(lm)
        sel = Window(np.tile([1.0, 0, 0, 0, 0, 0, 0], (n_kf, 1)), fx, [[1.0, 0, 0]], [[1.0, 0, 0, 0, 0, 0, 0]], np.zeros((n_lm, 3)),
                     np.ones(n_lm), np.zeros(n_lm + 1, dtype=np.int32), [], [], [], [], **scalars)
        res = Result(sel, 256)
        _check(lib().kba_track_solve(self._p, n_kf, kfp, fx.ctypes.data_as(C.POINTER(C.c_uint8)), n_lm, lmp, C.byref(sel.c),
                                     C.byref(opt or default_options()), C.byref(res.c)))
        return res

    def transfer_bytes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().kba_track_transfer_bytes(self._p, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def close(self):
        if self._p:
            lib().kba_track_destroy(self._p)
            self._p = C.c_void_p()


SHARD_ID_BYTES = 128


def shard_unique_id():
    """NCCL unique id (bytes) -- create on rank 0 and broadcast to the other ranks"""
    buf = C.create_string_buffer(SHARD_ID_BYTES)
    _check(lib().kba_shard_unique_id(buf))
    return bytes(buf.raw)


class ShardComm:
    """NCCL communicator of the sharded window solve (kba_shard_comm_create is collective over all ranks)"""

    def __init__(self, handle, rank, world, unique_id):
        assert len(unique_id) == SHARD_ID_BYTES
        self._p = C.c_void_p()
        self.rank, self.world = rank, world
        _check(lib().kba_shard_comm_create(handle._p, rank, world, C.c_char_p(unique_id), C.byref(self._p)))

    def close(self):
        if self._p:
            lib().kba_shard_comm_destroy(self._p)
            self._p = C.c_void_p()


class Handle:
    """One solver handle per host thread / GPU (kba_create)."""

    def __init__(self, device=0, stream=None):
        self._p = C.c_void_p()
        _check(lib().kba_create(C.byref(self._p), device))
        if stream is not None:
            _check(lib().kba_set_stream(self._p, C.c_void_p(stream)))

    def default_options(self):
        return default_options()

    def solve_window(self, win, opt=None, iterations_capacity=256):
        res = Result(win, iterations_capacity)
        _check(lib().kba_solve_window(self._p, C.byref(win.c), C.byref(opt or default_options()), C.byref(res.c)))
        return res

    def solve_batch(self, windows, opt=None, iterations_capacity=1):
        results = [Result(w, iterations_capacity) for w in windows]
        warr = (KbaWindow * len(windows))(*[w.c for w in windows])
        rarr = (KbaResult * len(windows))(*[r.c for r in results])
        _check(lib().kba_solve_batch(self._p, len(windows), warr, C.byref(opt or default_options()), rarr))
        for r, c in zip(results, rarr):
            r.c = c
        self._keep = rarr
        return results

    def evaluate(self, win, opt=None):
        n = max(win.n_obs, 1)
        r = np.zeros((n, 3)); jp = np.zeros((n, 3, 6)); jl = np.zeros((n, 3, 3)); cost = np.zeros(1)
        failed = np.zeros(1, dtype=np.int32)
        out = KbaEvalOut()
        out.residual = r.ctypes.data_as(c_double_p); out.jac_pose = jp.ctypes.data_as(c_double_p)
        out.jac_lm = jl.ctypes.data_as(c_double_p); out.cost = cost.ctypes.data_as(c_double_p)
        out.failed = failed.ctypes.data_as(c_int32_p)
        _check(lib().kba_eval(self._p, C.byref(win.c), C.byref(opt or default_options()), C.byref(out)))
        return r[:win.n_obs], jp[:win.n_obs], jl[:win.n_obs], float(cost[0]), int(failed[0])

    def batch(self, windows):
        return Batch(self, windows)

    def init_landmarks(self, win):
        """push() landmark initialisation for every landmark of `win` on the device: (positions, flags, device ms);
        flags bit 0 = created, bit 1 = in front of every observing camera"""
        pos = np.zeros((max(win.n_lm, 1), 3))
        flags = np.zeros(max(win.n_lm, 1), dtype=np.uint8)
        ms = C.c_float()
        _check(lib().kba_init_landmarks(self._p, C.byref(win.c), pos.ctypes.data_as(c_double_p),
                                        flags.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)))
        return pos[:win.n_lm], flags[:win.n_lm], ms.value

    def lidar_depth(self, cloud, T_cam_lidar, intr, features_uv, opt=None):
        """cloud [n, stride>=3] float32, features_uv [m, 2] float32 -> (depth [m] float32 (-1 = none), device ms)"""
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        feats = np.ascontiguousarray(features_uv, dtype=np.float32).reshape(-1, 2)
        T = np.ascontiguousarray(T_cam_lidar, dtype=np.float64); K = np.ascontiguousarray(intr, dtype=np.float64)
        out = np.zeros(max(len(feats), 1), dtype=np.float32)
        ms = C.c_float()
        fp = C.POINTER(C.c_float)
        _check(lib().kba_lidar_depth(self._p, cloud.ctypes.data_as(fp), cloud.shape[0], cloud.shape[1],
                                     T.ctypes.data_as(c_double_p), K.ctypes.data_as(c_double_p), feats.ctypes.data_as(fp),
                                     len(feats), C.byref(opt or lidar_default_options()), out.ctypes.data_as(fp),
                                     C.byref(ms)))
        return out[:len(feats)], ms.value

    def counters(self, reset=False):
        c = KbaCounters()
        _check(lib().kba_get_counters(self._p, C.byref(c), 1 if reset else 0))
        return c

    def enable_kernel_timing(self, on=True):
        _check(lib().kba_enable_kernel_timing(self._p, 1 if on else 0))

    def close(self):
        if self._p:
            lib().kba_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
