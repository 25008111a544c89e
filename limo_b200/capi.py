"""ctypes binding of the CUDA library limo_b200/libkba_b200.so (C ABI: include/kba_b200.h).

The product path: there is no CPU fallback.  Importing works without a GPU (symbols can be inspected), but every
computing call fails with KBA_ERR_CUDA when no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

from .capi_types import (KbaCounters, KbaEvalOut, KbaLidarOptions, KbaOptions, KbaResult, KbaWindow, Result, c_double_p, c_int32_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KBA_LIB_PATH") or os.path.join(_HERE, "libkba_b200.so")  # KBA_LIB_PATH: instrumented builds
_lib = None

SYMBOLS = ["kba_version", "kba_last_error", "kba_default_options", "kba_create", "kba_destroy", "kba_set_stream",
           "kba_solve_window", "kba_solve_batch", "kba_eval", "kba_batch_create", "kba_batch_upload",
           "kba_batch_solve", "kba_batch_download", "kba_batch_transfer_bytes", "kba_batch_jacobian_pass", "kba_batch_destroy",
           "kba_get_counters", "kba_enable_kernel_timing", "kba_lidar_default_options", "kba_lidar_depth",
           "kba_shard_unique_id", "kba_shard_comm_create", "kba_shard_comm_destroy", "kba_batch_set_shard",
           "kba_init_landmarks"]


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
