"""ctypes binding of include/elas_b200.h (harness side: tests, bench, smoke).

Fails loudly when the CUDA library is missing or no device is usable: there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ELAS_B200_LIB") or os.path.join(HERE, "libelas_b200.so")     # same variable as the drop-in elas.cpp

E_FEW_SUPPORT = 1

_PARAM_FIELDS = [
    ("disp_min", C.c_int32), ("disp_max", C.c_int32), ("support_threshold", C.c_float),
    ("support_texture", C.c_int32), ("candidate_stepsize", C.c_int32),
    ("incon_window_size", C.c_int32), ("incon_threshold", C.c_int32),
    ("incon_min_support", C.c_int32), ("add_corners", C.c_int32), ("grid_size", C.c_int32),
    ("beta", C.c_float), ("gamma", C.c_float), ("sigma", C.c_float), ("sradius", C.c_float),
    ("match_texture", C.c_int32), ("lr_threshold", C.c_int32),
    ("speckle_sim_threshold", C.c_float), ("speckle_size", C.c_int32),
    ("ipol_gap_width", C.c_int32), ("filter_median", C.c_int32),
    ("filter_adaptive_mean", C.c_int32), ("postprocess_only_left", C.c_int32),
    ("subsampling", C.c_int32),
]


class Params(C.Structure):
    """elas_b200_params: POD mirror of Elas::parameters (reference elas.h:59-85)."""
    _fields_ = _PARAM_FIELDS

    def copy(self, **overrides):
        q = Params.from_buffer_copy(bytes(self))
        for k, v in overrides.items():
            setattr(q, k, v)
        return q


class View(C.Structure):
    """elas_b200_view: intrinsics, range limit, gain and pose used by StereoThread::createCurrentMap
    (reference stereothread.cpp:180-255, :441-447)."""
    _fields_ = [("f", C.c_float), ("cu", C.c_float), ("cv", C.c_float), ("base", C.c_float),
                ("max_dist", C.c_float), ("gain", C.c_float), ("H", C.c_double * 12)]


class Map3D(C.Structure):
    """elas_b200_map3d: the five planes of StereoThread::map3d (stereothread.h:51-60)."""
    _fields_ = [("I", C.c_void_p), ("D", C.c_void_p), ("X", C.c_void_p), ("Y", C.c_void_p), ("Z", C.c_void_p)]


def _image(a):
    """A uint8 image as the C ABI takes it: unit column stride, any row pitch >= width (bytes_per_line = strides[0]);
    views into wider rows are passed as they are, anything else is copied."""
    if isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1 and a.strides[0] >= a.shape[1]:
        return a
    return np.ascontiguousarray(a, np.uint8)


class LibraryMissing(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "elas_b200_default_params", "elas_b200_stereomapper_params", "elas_b200_process",
    "elas_b200_create", "elas_b200_create_ex", "elas_b200_create_grouped", "elas_b200_frames_per_group",
    "elas_b200_mesh_on_device", "elas_b200_time_matching_ex", "elas_b200_destroy", "elas_b200_multi_create",
    "elas_b200_fuse", "elas_b200_matcher_filters", "elas_b200_multi_destroy", "elas_b200_multi_device_count", "elas_b200_multi_context", "elas_b200_multi_process_batch", "elas_b200_process_ctx", "elas_b200_process_batch",
    "elas_b200_process_batch_device", "elas_b200_stage_capture", "elas_b200_stage_bytes",
    "elas_b200_stage_read", "elas_b200_host_stage", "elas_b200_launch_count",
    "elas_b200_stage_timing", "elas_b200_stage_times", "elas_b200_host_times", "elas_b200_time_matching",
    "elas_b200_version", "elas_b200_device_count", "elas_b200_colormap", "elas_b200_reproject",
    "elas_b200_time_view",
]


def build_library():
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE])


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    P, u8p, f32p, i32p = C.POINTER(Params), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)
    lib.elas_b200_default_params.argtypes = [P, C.c_int32]
    lib.elas_b200_default_params.restype = None
    lib.elas_b200_stereomapper_params.argtypes = [P]
    lib.elas_b200_stereomapper_params.restype = None
    lib.elas_b200_process.argtypes = [P, u8p, u8p, f32p, f32p, i32p]
    lib.elas_b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, P, C.c_int32, C.c_int32, C.c_int32]
    lib.elas_b200_create_ex.argtypes = [C.POINTER(C.c_void_p), C.c_int32, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.elas_b200_create_grouped.argtypes = [C.POINTER(C.c_void_p), C.c_int32, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.elas_b200_frames_per_group.argtypes = [C.c_void_p]
    lib.elas_b200_mesh_on_device.argtypes = [C.c_void_p]
    lib.elas_b200_time_matching_ex.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, i32p]
    lib.elas_b200_time_matching_ex.restype = C.c_float
    lib.elas_b200_destroy.argtypes = [C.c_void_p]
    lib.elas_b200_destroy.restype = None
    lib.elas_b200_multi_create.argtypes = [C.POINTER(C.c_void_p), i32p, C.c_int32, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.elas_b200_multi_destroy.argtypes = [C.c_void_p]
    lib.elas_b200_multi_destroy.restype = None
    lib.elas_b200_multi_device_count.argtypes = [C.c_void_p]
    lib.elas_b200_multi_context.argtypes = [C.c_void_p, C.c_int32]
    lib.elas_b200_multi_context.restype = C.c_void_p
    lib.elas_b200_multi_process_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, i32p]
    lib.elas_b200_process_ctx.argtypes = [C.c_void_p, C.c_int32, u8p, u8p, f32p, f32p, C.c_int32]
    for fn in (lib.elas_b200_process_batch, lib.elas_b200_process_batch_device):
        fn.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, i32p]
    lib.elas_b200_stage_capture.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.elas_b200_stage_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_char_p]
    lib.elas_b200_stage_bytes.restype = C.c_int64
    lib.elas_b200_stage_read.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_int64]
    lib.elas_b200_host_stage.argtypes = [P, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, i32p]
    lib.elas_b200_launch_count.argtypes = [C.c_void_p]
    lib.elas_b200_launch_count.restype = C.c_int64
    lib.elas_b200_stage_timing.argtypes = [C.c_void_p, C.c_int32]
    lib.elas_b200_stage_times.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int32]
    lib.elas_b200_host_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
    lib.elas_b200_time_matching.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.elas_b200_time_matching.restype = C.c_float
    lib.elas_b200_colormap.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.elas_b200_reproject.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(View)] + [C.c_void_p] * 5
    lib.elas_b200_time_view.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    lib.elas_b200_fuse.argtypes = [C.c_void_p, C.c_int32, C.POINTER(View), C.POINTER(Map3D), C.POINTER(Map3D),
                                   C.c_void_p, i32p, C.c_void_p, i32p]
    lib.elas_b200_matcher_filters.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.POINTER(C.c_float)]
    lib.elas_b200_version.restype = C.c_char_p
    lib.elas_b200_device_count.restype = C.c_int32
    _lib = lib
    return lib


def _preset(which):
    p = Params()
    load_library().elas_b200_default_params(C.byref(p), which)
    return p


def robotics():
    return _preset(0)


def middlebury():
    return _preset(1)


def stereomapper(dmax=255):
    p = Params()
    load_library().elas_b200_stereomapper_params(C.byref(p))
    p.disp_max = dmax
    return p


def demo(dmax=255):
    return robotics().copy(postprocess_only_left=0, disp_max=dmax)


_STAGE_DTYPES = {
    "desc1": np.uint8, "desc2": np.uint8, "dcan_raw": np.int16, "dcan_incon": np.int16,
    "dcan": np.int16, "lattice_dims": np.int32, "support": np.int32, "tri1": np.int32,
    "tri2": np.int32, "planes1": np.float32, "planes2": np.float32, "grid1": np.int32,
    "grid2": np.int32, "grid_dims": np.int32, "grid1_bits": np.uint32, "grid2_bits": np.uint32,
}


def process(I1, I2, params):
    """The synchronous drop-in call elas_b200_process (what Elas::process binds)."""
    lib = load_library()
    I1, I2 = _image(I1), _image(I2)
    if I1.strides[0] != I2.strides[0]:
        I1, I2 = np.ascontiguousarray(I1), np.ascontiguousarray(I2)
    H, W = I1.shape
    shape = (H // 2, W // 2) if params.subsampling else (H, W)
    D1 = np.full(shape, -77.0, np.float32)
    D2 = np.full(shape, -77.0, np.float32)
    dims = (C.c_int32 * 3)(W, H, I1.strides[0])
    rc = lib.elas_b200_process(C.byref(params), I1.ctypes.data, I2.ctypes.data, D1.ctypes.data, D2.ctypes.data, dims)
    if rc < 0:
        raise RuntimeError(f"elas_b200_process failed with {rc}")
    return rc, D1, D2


def host_stage(params, width, height, dcan):
    """elas_b200_host_stage: lattice filters + support list + Delaunay + planes on the CPU."""
    lib = load_library()
    dcan = np.ascontiguousarray(dcan, np.int16).copy()
    cap = dcan.size + 6
    tcap = 2 * cap + 8
    sup = np.zeros((cap, 3), np.int32)
    tri1 = np.zeros((tcap, 3), np.int32)
    tri2 = np.zeros((tcap, 3), np.int32)
    pl1 = np.zeros((tcap, 6), np.float32)
    pl2 = np.zeros((tcap, 6), np.float32)
    n = (C.c_int32 * 3)()
    rc = lib.elas_b200_host_stage(C.byref(params), width, height, dcan.ctypes.data, sup.ctypes.data, cap,
                                  tri1.ctypes.data, tri2.ctypes.data, pl1.ctypes.data, pl2.ctypes.data, tcap, n)
    if rc < 0:
        raise RuntimeError(f"elas_b200_host_stage failed with {rc}")
    return {"rc": rc, "dcan": dcan, "support": sup[:n[0]].copy(), "tri1": tri1[:n[1]].copy(),
            "tri2": tri2[:n[2]].copy(), "planes1": pl1[:n[1]].copy(), "planes2": pl2[:n[2]].copy()}


class ElasB200:
    """A persistent context: n_slots frame groups of frames_per_group frames each in flight on one device
    (frames_per_group=1: elas_b200_create_ex, the "slot" model; 0 = chosen from the frame size)."""

    def __init__(self, params, width, height, n_slots=1, device=0, n_workers=0, frames_per_group=1):
        self.lib = load_library()
        if self.lib.elas_b200_device_count() < 1:
            raise RuntimeError("elas_b200: no CUDA device; this library has no CPU fallback")
        self.params, self.W, self.H, self.n_slots, self.device = params, width, height, n_slots, device
        self.shape = (height // 2, width // 2) if params.subsampling else (height, width)
        self.ctx = C.c_void_p()
        rc = self.lib.elas_b200_create_grouped(C.byref(self.ctx), device, C.byref(params), width, height, n_slots,
                                               frames_per_group, n_workers)
        if rc != 0:
            raise RuntimeError(f"elas_b200_create_grouped failed with {rc}")
        self.frames_per_group = int(self.lib.elas_b200_frames_per_group(self.ctx))
        self.mesh_on_device = bool(self.lib.elas_b200_mesh_on_device(self.ctx))

    def close(self):
        if self.ctx:
            self.lib.elas_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, I1, I2, slot=0, capture=False):
        I1, I2 = _image(I1), _image(I2)
        if I1.strides[0] != I2.strides[0]:
            I1, I2 = np.ascontiguousarray(I1), np.ascontiguousarray(I2)
        assert I1.shape == (self.H, self.W) and I2.shape == I1.shape
        D1 = np.full(self.shape, -77.0, np.float32)
        D2 = np.full(self.shape, -77.0, np.float32)
        self.lib.elas_b200_stage_capture(self.ctx, slot, 1 if capture else 0)
        rc = self.lib.elas_b200_process_ctx(self.ctx, slot, I1.ctypes.data, I2.ctypes.data, D1.ctypes.data,
                                            D2.ctypes.data, I1.strides[0])
        if rc < 0:
            raise RuntimeError(f"elas_b200_process_ctx failed with {rc}")
        return rc, D1, D2

    def colormap(self, D1=None, slot=0):
        """HSV colour map (stereothread.cpp:116-147) of D1, or of the map the slot's last frame left in HBM."""
        if D1 is not None:
            D1 = np.ascontiguousarray(D1, np.float32)
            assert D1.shape == self.shape
        out = np.empty(self.shape + (3,), np.float32)
        rc = self.lib.elas_b200_colormap(self.ctx, slot, D1.ctypes.data if D1 is not None else None, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"elas_b200_colormap failed with {rc}")
        return out

    def reproject(self, view, H, I1=None, D1=None, slot=0):
        """StereoThread::createCurrentMap (stereothread.cpp:180-255): returns I, D, X, Y, Z.
        view = (f, cu, cv, base, max_dist, gain), H = 3x4 pose; I1/D1 None = the slot's last frame."""
        v = View(*[float(x) for x in view], (C.c_double * 12)(*[float(x) for x in np.asarray(H, np.float64).reshape(12)]))
        pitch = 0
        if I1 is not None:
            assert I1.dtype == np.uint8 and I1.shape == (self.H, self.W) and I1.strides[1] == 1
            pitch = I1.strides[0]
        if D1 is not None:
            D1 = np.ascontiguousarray(D1, np.float32)
        outs = [np.empty((self.H, self.W), np.float32) for _ in range(5)]
        rc = self.lib.elas_b200_reproject(self.ctx, slot, I1.ctypes.data if I1 is not None else None, pitch,
                                          D1.ctypes.data if D1 is not None else None, C.byref(v),
                                          *[o.ctypes.data for o in outs])
        if rc != 0:
            raise RuntimeError(f"elas_b200_reproject failed with {rc}")
        return outs

    def fuse(self, view, H, cur, prev=None, slot=0):
        """StereoThread::addDisparityMapToReconstruction (stereothread.cpp:290-437).  cur / prev = [I, D, X, Y, Z]
        (prev: what the previous call returned, or None).  Returns (fused current map, previous D after the call or
        None, points_prev, points_curr)."""
        v = View(*[float(x) for x in view], (C.c_double * 12)(*[float(x) for x in np.asarray(H, np.float64).reshape(12)]))
        cur = [np.ascontiguousarray(a, np.float32).copy() for a in cur]
        pv = [np.ascontiguousarray(a, np.float32).copy() for a in prev] if prev is not None else None
        n = self.W * self.H
        pts_prev, pts_curr = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        n_prev, n_curr = C.c_int32(0), C.c_int32(0)
        cm = Map3D(*[a.ctypes.data for a in cur])
        pm = Map3D(*[a.ctypes.data for a in pv]) if pv else None
        rc = self.lib.elas_b200_fuse(self.ctx, slot, C.byref(v), C.byref(pm) if pm else None, C.byref(cm),
                                     pts_prev.ctypes.data, C.byref(n_prev), pts_curr.ctypes.data, C.byref(n_curr))
        if rc != 0:
            raise RuntimeError(f"elas_b200_fuse failed with {rc}")
        return cur, (pv[1] if pv else None), pts_prev[:n_prev.value].copy(), pts_curr[:n_curr.value].copy()

    def time_view(self, iters=20, slot=0):
        """Mean ms per launch of (k_colormap, k_reproject) on the slot's device-resident last frame."""
        ms = (C.c_float * 2)()
        rc = self.lib.elas_b200_time_view(self.ctx, slot, iters, ms)
        if rc != 0:
            raise RuntimeError(f"elas_b200_time_view failed with {rc}")
        return float(ms[0]), float(ms[1])

    def stage(self, name, slot=0):
        n = self.lib.elas_b200_stage_bytes(self.ctx, slot, name.encode())
        if n < 0:
            return None
        dt = np.dtype(_STAGE_DTYPES.get(name, np.float32))
        a = np.empty(n // dt.itemsize, dt)
        rc = self.lib.elas_b200_stage_read(self.ctx, slot, name.encode(), a.ctypes.data, n)
        assert rc == 0
        return a

    @staticmethod
    def _ptr_array(ptrs):
        return (C.c_void_p * len(ptrs))(*ptrs)

    def process_batch_ptrs(self, I1, I2, D1, D2, bytes_per_line, device=False):
        """Raw pointers (ints) of n frames: host (pinned or pageable) or device memory."""
        n = len(I1)
        status = (C.c_int32 * n)()
        fn = self.lib.elas_b200_process_batch_device if device else self.lib.elas_b200_process_batch
        rc = fn(self.ctx, n, self._ptr_array(I1), self._ptr_array(I2), self._ptr_array(D1),
                self._ptr_array(D2), bytes_per_line, status)
        if rc < 0:
            raise RuntimeError(f"elas_b200_process_batch failed with {rc}")
        return list(status)

    def process_batch(self, lefts, rights):
        lefts = [_image(a) for a in lefts]
        rights = [_image(a) for a in rights]
        if len({a.strides[0] for a in lefts + rights}) != 1:          # one bytes_per_line per call
            lefts = [np.ascontiguousarray(a) for a in lefts]; rights = [np.ascontiguousarray(a) for a in rights]
        D1 = [np.full(self.shape, -77.0, np.float32) for _ in lefts]
        D2 = [np.full(self.shape, -77.0, np.float32) for _ in lefts]
        status = self.process_batch_ptrs([a.ctypes.data for a in lefts], [a.ctypes.data for a in rights],
                                         [a.ctypes.data for a in D1], [a.ctypes.data for a in D2],
                                         lefts[0].strides[0])
        return status, D1, D2

    def launch_count(self):
        return int(self.lib.elas_b200_launch_count(self.ctx))

    def set_timing(self, on):
        self.lib.elas_b200_stage_timing(self.ctx, 1 if on else 0)

    def stage_times(self, slot=0):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = self.lib.elas_b200_stage_times(self.ctx, slot, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(max(n, 0))]

    def host_times(self, reset=True):
        """Per-frame host wall time by phase (ms): submit A, wait A, host stage, submit B, wait B."""
        ms = (C.c_double * 5)()
        frames = C.c_int64(0)
        self.lib.elas_b200_host_times(self.ctx, ms, C.byref(frames), 1 if reset else 0)
        n = max(frames.value, 1)
        names = ["submit", "wait", "host_stage", "finish", "unused"]
        return {k: ms[i] / n for i, k in enumerate(names)}, frames.value

    def time_matching(self, iters=20, flush_l2=True, slot=0, per_frame=True):
        """Mean ms of the matching kernel on the tables of the slot's last launch chain: per frame (default) or
        per launch with the number of frames it processes."""
        frames = C.c_int32(0)
        ms = float(self.lib.elas_b200_time_matching_ex(self.ctx, slot, iters, 1 if flush_l2 else 0, C.byref(frames)))
        if ms < 0:
            raise RuntimeError("elas_b200_time_matching failed (run a frame through the slot first)")
        return ms / frames.value if per_frame else (ms, frames.value)


class ElasB200Multi:
    """elas_b200_multi_*: one context per device in ONE process, frame i of a batch -> device i mod n."""

    def __init__(self, params, width, height, devices, n_groups=2, frames_per_group=0, n_workers=0):
        self.lib = load_library()
        self.shape = (height // 2, width // 2) if params.subsampling else (height, width)
        self.handle = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices)
        rc = self.lib.elas_b200_multi_create(C.byref(self.handle), devs, len(devices), C.byref(params), width, height,
                                             n_groups, frames_per_group, n_workers)
        if rc != 0:
            raise RuntimeError(f"elas_b200_multi_create failed with {rc}")

    def close(self):
        if self.handle:
            self.lib.elas_b200_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def process_batch(self, lefts, rights):
        lefts = [np.ascontiguousarray(a, np.uint8) for a in lefts]
        rights = [np.ascontiguousarray(a, np.uint8) for a in rights]
        n = len(lefts)
        D1 = [np.full(self.shape, -77.0, np.float32) for _ in lefts]
        D2 = [np.full(self.shape, -77.0, np.float32) for _ in lefts]
        arr = lambda xs: (C.c_void_p * n)(*[a.ctypes.data for a in xs])
        status = (C.c_int32 * n)()
        rc = self.lib.elas_b200_multi_process_batch(self.handle, n, arr(lefts), arr(rights), arr(D1), arr(D2),
                                                    lefts[0].strides[0], status)
        if rc < 0:
            raise RuntimeError(f"elas_b200_multi_process_batch failed with {rc}")
        return list(status), D1, D2


def matcher_filters(I, device=0, iters=1, timing=False):
    """The feature filters of libviso2's Matcher (filter::sobel5x5 / blob5x5 / checkerboard5x5 as called at
    matcher.cpp:799-801) on a uint8 image [h][bytes_per_line]; returns (du, dv, f1, f2) [+ ms per pass]."""
    lib = load_library()
    I = np.ascontiguousarray(I, np.uint8)
    h, w = I.shape
    du, dv = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
    f1, f2 = np.empty((h, w), np.int16), np.empty((h, w), np.int16)
    ms = C.c_float(0)
    rc = lib.elas_b200_matcher_filters(device, I.ctypes.data, w, h, du.ctypes.data, dv.ctypes.data, f1.ctypes.data,
                                       f2.ctypes.data, iters, C.byref(ms) if timing else None)
    if rc != 0:
        raise RuntimeError(f"elas_b200_matcher_filters failed with {rc}")
    return (du, dv, f1, f2, ms.value) if timing else (du, dv, f1, f2)
