"""TEST INFRASTRUCTURE -- ctypes bindings of the two CPU checkers.

  RefElas     oracle/_ref/libelas_ref.so    the unmodified reference libelas + stage-dump harness
  OracleElas  oracle/_build/libelas_oracle.so  the plain-C restatement (elas_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (stereo-vision_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libelas_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libelas_oracle.so")
VIEW_REF_SO = os.path.join(HERE, "_ref", "libview_ref.so")
FILTER_REF_SO = os.path.join(HERE, "_ref", "libvisofilter_ref.so")

PARAM_FIELDS = [
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
    """POD mirror of Elas::parameters (include/elas_b200.h; reference elas.h:59-85)."""
    _fields_ = PARAM_FIELDS

    def copy(self, **overrides):
        q = Params.from_buffer_copy(bytes(self))
        for k, v in overrides.items():
            setattr(q, k, v)
        return q


def robotics():
    """Elas::parameters(ROBOTICS), elas.h:93-118."""
    return Params(0, 255, 0.85, 10, 5, 5, 5, 5, 0, 20, 0.02, 3, 1, 2, 1, 2, 1, 200, 3, 0, 1, 1, 0)


def middlebury():
    """Elas::parameters(MIDDLEBURY), elas.h:121-146."""
    return Params(0, 255, 0.95, 10, 5, 5, 5, 5, 1, 20, 0.02, 5, 1, 3, 0, 2, 1, 200, 5000, 1, 0, 0, 0)


def stereomapper(dmax=255):
    """What StereoThread::run sets (stereothread.cpp:76-80)."""
    return robotics().copy(postprocess_only_left=1, filter_adaptive_mean=1, support_texture=30,
                           disp_max=dmax)


def demo(dmax=255):
    """What libelas/src/main.cpp:61-62 sets."""
    return robotics().copy(postprocess_only_left=0, disp_max=dmax)


STAGE_DTYPES = {
    "desc1": np.uint8, "desc2": np.uint8, "dcan_raw": np.int16, "dcan_incon": np.int16,
    "dcan": np.int16, "lattice_dims": np.int32, "support": np.int32, "tri1": np.int32,
    "tri2": np.int32, "planes1": np.float32, "planes2": np.float32, "grid1": np.int32,
    "grid2": np.int32, "grid_dims": np.int32,
}


def stage_dtype(name):
    return STAGE_DTYPES.get(name, np.float32)


def build(target="all"):
    subprocess.check_call(["make", "-s", "-C", HERE, target])


def have_ref():
    return os.path.exists(REF_SO)


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class _StageLib:
    """Shared shape of both checkers: <prefix>_process, _run_stages, _stage_bytes, _stage_read."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        for fn in ("process", "run_stages"):
            f = getattr(self.lib, f"{prefix}_{fn}")
            f.restype = C.c_int32
            f.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        sb = getattr(self.lib, f"{prefix}_stage_bytes")
        sb.restype = C.c_int64
        sb.argtypes = [C.c_char_p]
        sr = getattr(self.lib, f"{prefix}_stage_read")
        sr.restype = C.c_int32
        sr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]

    @staticmethod
    def _prep(I1, I2, p):
        I1 = np.ascontiguousarray(I1, np.uint8)
        I2 = np.ascontiguousarray(I2, np.uint8)
        H, W = I1.shape
        dims = (C.c_int32 * 3)(W, H, I1.strides[0])
        shape = (H // 2, W // 2) if p.subsampling else (H, W)
        # sentinel so the "fewer than 3 support points" early return (elas.cpp:69-75) is visible
        D1 = np.full(shape, -77.0, np.float32)
        D2 = np.full(shape, -77.0, np.float32)
        return I1, I2, dims, D1, D2

    def process(self, I1, I2, p):
        I1, I2, dims, D1, D2 = self._prep(I1, I2, p)
        rc = getattr(self.lib, f"{self.prefix}_process")(C.byref(p), _u8(I1), _u8(I2), _f32(D1), _f32(D2), dims)
        return rc, D1, D2

    def run_stages(self, I1, I2, p, names=None):
        """Runs the pipeline stage by stage; returns (rc, D1, D2, {stage name: flat ndarray})."""
        I1, I2, dims, D1, D2 = self._prep(I1, I2, p)
        rc = getattr(self.lib, f"{self.prefix}_run_stages")(C.byref(p), _u8(I1), _u8(I2), _f32(D1), _f32(D2), dims)
        out = {}
        wanted = names or list(STAGE_DTYPES) + [f"D{i}_{s}" for i in (1, 2) for s in ("raw", "lr", "seg", "gap", "mean")] + ["D1", "D2"]
        for name in wanted:
            a = self.stage(name)
            if a is not None:
                out[name] = a
        return rc, D1, D2, out

    def stage(self, name):
        n = getattr(self.lib, f"{self.prefix}_stage_bytes")(name.encode())
        if n < 0:
            return None
        dt = np.dtype(stage_dtype(name))
        a = np.empty(n // dt.itemsize, dt)
        rc = getattr(self.lib, f"{self.prefix}_stage_read")(name.encode(), a.ctypes.data, n)
        assert rc == 0
        return a


class RefElas(_StageLib):
    def __init__(self):
        super().__init__(REF_SO, "ref")
        self.lib.ref_time_process.restype = C.c_double
        self.lib.ref_time_process.argtypes = [
            C.POINTER(Params), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_float),
            C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double)]
        self.lib.ref_delaunay.restype = C.c_int32
        self.lib.ref_delaunay.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]

    def time_process(self, I1, I2, p, reps=3):
        I1, I2, dims, D1, D2 = self._prep(I1, I2, p)
        mean = C.c_double(0)
        best = self.lib.ref_time_process(C.byref(p), _u8(I1), _u8(I2), _f32(D1), _f32(D2), dims, reps, C.byref(mean))
        return best, mean.value

    def delaunay(self, support, right_image):
        s = np.ascontiguousarray(support, np.int32).reshape(-1, 3)
        out = np.empty((2 * len(s) + 8, 3), np.int32)
        n = self.lib.ref_delaunay(s.ctypes.data, len(s), int(right_image), out.ctypes.data, len(out))
        return out[:n].copy()


class OracleElas(_StageLib):
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build("oracle")
        super().__init__(ORACLE_SO, "oracle")
        self.lib.oracle_delaunay.restype = C.c_int32
        self.lib.oracle_delaunay.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]

    def delaunay(self, support, right_image):
        s = np.ascontiguousarray(support, np.int32).reshape(-1, 3)
        out = np.empty((2 * len(s) + 8, 3), np.int32)
        n = self.lib.oracle_delaunay(s.ctypes.data, len(s), int(right_image), out.ctypes.data, len(out))
        return out[:n].copy()


class ViewChecker:
    """Colour map + back-projection (stereothread.cpp:116-147, :180-255): the plain-C restatement
    (kind="oracle") or the reference's own statements compiled by oracle/Makefile (kind="ref")."""

    def __init__(self, kind="oracle"):
        self.kind = kind
        if kind == "ref":
            self.lib, prefix = C.CDLL(VIEW_REF_SO), "ref"
        else:
            if not os.path.exists(ORACLE_SO):
                build("oracle")
            self.lib, prefix = C.CDLL(ORACLE_SO), "oracle"
        self._colormap = getattr(self.lib, f"{prefix}_colormap")
        self._colormap.restype = None
        self._colormap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        self._reproject = getattr(self.lib, f"{prefix}_reproject")
        self._reproject.restype = None
        self._reproject.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p] + [C.c_void_p] * 5

    def fuse(self, I1, D1, view, H, prev=None):
        """StereoThread::addDisparityMapToReconstruction (stereothread.cpp:290-437) for one frame: builds the current
        map (createCurrentMap) and fuses it with `prev` (the fused map (I, D, X, Y, Z) the previous call returned, or
        None).  Returns (fused current map [I, D, X, Y, Z], previous D after the call or None, points_prev, points_curr)."""
        I1 = np.ascontiguousarray(I1, np.uint8) if I1.strides[1] != 1 else I1
        D1 = np.ascontiguousarray(D1, np.float32)
        h, w = D1.shape
        view = np.ascontiguousarray(view, np.float32)
        H = np.ascontiguousarray(H, np.float64).reshape(12)
        n = w * h
        cur = [np.zeros((h, w), np.float32) for _ in range(5)]
        pts_prev, pts_curr = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        n_prev, n_curr = C.c_int32(0), C.c_int32(0)
        pv = [np.ascontiguousarray(a, np.float32).copy() for a in prev] if prev is not None else None
        if self.kind == "ref":
            P5 = C.c_void_p * 5
            parr = P5(*[a.ctypes.data for a in pv]) if pv else P5(None, None, None, None, None)
            carr = P5(*[a.ctypes.data for a in cur])
            self.lib.ref_fuse.restype = None
            self.lib.ref_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
            self.lib.ref_fuse(I1.ctypes.data, D1.ctypes.data, w, h, I1.strides[0], view.ctypes.data, H.ctypes.data,
                              parr, carr, pts_prev.ctypes.data, C.byref(n_prev), pts_curr.ctypes.data, C.byref(n_curr))
        else:
            cur = list(self.reproject(I1, D1, view, H))
            self.lib.oracle_fuse.restype = None
            self.lib.oracle_fuse.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p] + [C.c_void_p] * 10 + \
                                            [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int32)]
            pp = [a.ctypes.data for a in pv] if pv else [None] * 5
            self.lib.oracle_fuse(w, h, view.ctypes.data, H.ctypes.data, *pp, *[a.ctypes.data for a in cur],
                                 pts_prev.ctypes.data, C.byref(n_prev), pts_curr.ctypes.data, C.byref(n_curr))
        return cur, (pv[1] if pv else None), pts_prev[:n_prev.value].copy(), pts_curr[:n_curr.value].copy()

    def colormap(self, D1):
        D1 = np.ascontiguousarray(D1, np.float32)
        out = np.empty(D1.shape + (3,), np.float32)
        self._colormap(D1.ctypes.data, D1.shape[1], D1.shape[0], out.ctypes.data)
        return out

    def reproject(self, I1, D1, view, H):
        """view = (f, cu, cv, base, max_dist, gain); H = 3x4 pose.  Returns I, D, X, Y, Z."""
        I1 = np.ascontiguousarray(I1, np.uint8)
        D1 = np.ascontiguousarray(D1, np.float32)
        h, w = D1.shape
        view = np.ascontiguousarray(view, np.float32)
        H = np.ascontiguousarray(H, np.float64).reshape(12)
        outs = [np.empty((h, w), np.float32) for _ in range(5)]
        self._reproject(I1.ctypes.data, D1.ctypes.data, w, h, I1.strides[0], view.ctypes.data, H.ctypes.data,
                        *[o.ctypes.data for o in outs])
        return outs


class MatcherFilterChecker:
    """The feature filters of libviso2's Matcher (filter.cpp:474-530 as called at matcher.cpp:799-801): the plain-C
    restatement (kind="oracle") or libviso2/src/filter.cpp itself compiled by oracle/Makefile (kind="ref")."""

    def __init__(self, kind="oracle"):
        if kind == "ref":
            self.fn = C.CDLL(FILTER_REF_SO).ref_matcher_filters
        else:
            if not os.path.exists(ORACLE_SO):
                build("oracle")
            self.fn = C.CDLL(ORACLE_SO).oracle_matcher_filters
        self.fn.restype = None
        self.fn.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 4

    def __call__(self, I):
        """I: uint8 [h][w], w a multiple of 16 (= bytes per line).  Returns (du, dv, f1, f2)."""
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        assert w % 16 == 0
        du, dv = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
        f1, f2 = np.zeros((h, w), np.int16), np.zeros((h, w), np.int16)
        self.fn(I.ctypes.data, w, h, du.ctypes.data, dv.ctypes.data, f1.ctypes.data, f2.ctypes.data)
        return du, dv, f1, f2


def have_filter_ref():
    return os.path.exists(FILTER_REF_SO)


def have_view_ref():
    return os.path.exists(VIEW_REF_SO)
