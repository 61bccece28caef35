"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): bit-exact for the integer support grid (descriptors, candidate
lattice, support points, triangles, candidate grid); disparity maps within +-1 level on pixels valid
in both with identical validity masks.  The kernels restate the reference's float expressions
operation by operation, so the float stages are asserted bit-exact as well; the +-1 bound is what a
failure report falls back to.
"""
import os

import numpy as np
import pytest

import checkers
import elas_b200
import synth
from helpers import (bits_equal, disparity_report, full_golden_cases, golden_cases, load_full_golden,
                     load_golden)

pytestmark = pytest.mark.gpu

INT_STAGES = ["desc1", "desc2", "dcan_raw", "dcan", "support", "tri1", "tri2", "grid1", "grid2"]
MAP_STAGES = ["D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D2_seg", "D1_gap", "D2_gap",
              "D1_mean", "D2_mean", "D1", "D2"]


def as_product_params(p):
    return elas_b200.Params.from_buffer_copy(bytes(p))


def run_cuda(L, R, p, capture=True, n_slots=1):
    H, W = L.shape
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=n_slots)
    try:
        rc, D1, D2 = e.process(L, R, capture=capture)
        stages = {}
        if capture:
            for k in INT_STAGES + MAP_STAGES + ["planes1", "planes2"]:
                a = e.stage(k)
                if a is not None:
                    stages[k] = a
        return rc, D1, D2, stages
    finally:
        e.close()


def check_against(st_cuda, st_cpu, D1, D2, tag):
    for k in INT_STAGES:
        assert k in st_cuda, f"{tag}: CUDA path did not produce stage {k}"
        assert bits_equal(st_cuda[k], st_cpu[k]), \
            f"{tag}: integer stage {k} not bit-exact ({int((st_cuda[k] != st_cpu[k]).sum()) if st_cuda[k].shape == st_cpu[k].shape else 'shape'} differ)"
    for k in ("planes1", "planes2"):
        assert bits_equal(st_cuda[k], st_cpu[k]), f"{tag}: {k} differ"
    for k in MAP_STAGES:
        rep = disparity_report(st_cuda[k], st_cpu[k], tol=1.0)
        assert rep["mask_mismatch"] == 0 and rep["over_tol"] == 0, f"{tag}: stage {k} outside +-1: {rep}"
        assert bits_equal(st_cuda[k], st_cpu[k]), f"{tag}: stage {k} within +-1 but not bit-exact: {rep}"
    assert bits_equal(D1.ravel(), st_cpu["D1"]) and bits_equal(D2.ravel(), st_cpu["D2"])


@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_reference_golden(name):
    """Committed outputs of the unmodified reference (tests/golden/make_golden.py)."""
    L, R, p, g = load_golden(name)
    rc, D1, D2, st = run_cuda(L, R, p)
    assert rc == 0
    for k in ["dcan_raw", "dcan", "support", "tri1", "tri2", "planes1", "planes2"]:
        assert bits_equal(st[k], g[k]), f"{name}: {k}"
    for k in ["D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D1_gap", "D1", "D2"]:
        rep = disparity_report(st[k], g[k])
        assert rep["mask_mismatch"] == 0 and rep["over_tol"] == 0, f"{name}: {k}: {rep}"
        assert bits_equal(st[k], g[k]), f"{name}: {k} not bit-exact: {rep}"


@pytest.mark.parametrize("name", full_golden_cases())
def test_cuda_matches_reference_on_full_size_real_pairs(name):
    """libelas/img/urban1..4 at full size (1344x391; SURVEY 8(d): the KITTI stand-ins): committed outputs
    of the unmodified reference.  Integer stages bit-exact incl. order; maps within +-1 and bit-exact."""
    L, R, p, g = load_full_golden(name)
    rc, D1, D2, st = run_cuda(L, R, p)
    assert rc == 0
    for k in ["dcan", "support", "tri1", "tri2"]:
        assert bits_equal(st[k], g[k].ravel()), f"{name}: {k}"
    for k, got in (("D1_raw", st["D1_raw"]), ("D2_raw", st["D2_raw"]), ("D1", D1.ravel()), ("D2", D2.ravel())):
        rep = disparity_report(got, g[k])
        assert rep["mask_mismatch"] == 0 and rep["over_tol"] == 0, f"{name}: {k}: {rep}"
        assert bits_equal(got, g[k].ravel()), f"{name}: {k} not bit-exact: {rep}"


CASES = [
    ("stereomapper-416x200", 416, 200, 95, 0, lambda d: checkers.stereomapper(d)),
    ("demo-416x200", 416, 200, 95, 1, lambda d: checkers.demo(d)),
    ("ragged-333x131", 333, 131, 63, 2, lambda d: checkers.stereomapper(d)),
    ("dmin3", 320, 160, 63, 6, lambda d: checkers.stereomapper(d).copy(disp_min=3)),
    ("dmax-not-multiple-of-32", 400, 180, 100, 9, lambda d: checkers.demo(d)),
    ("subsampling", 416, 200, 95, 3, lambda d: checkers.stereomapper(d).copy(subsampling=1)),
    ("subsampling-demo-odd-size", 417, 201, 95, 13, lambda d: checkers.demo(d).copy(subsampling=1)),
    ("middlebury-preset", 320, 160, 63, 4, lambda d: checkers.middlebury().copy(disp_max=d)),
    ("robotics+median", 320, 160, 63, 5, lambda d: checkers.demo(d).copy(filter_median=1)),
    ("wide-4096x160-d256-ten-segments", 4096, 160, 256, 12, lambda d: checkers.stereomapper(d)),
    ("K-1242x375-d255", 1242, 375, 255, 0, lambda d: checkers.stereomapper(d)),
    ("K-1242x375-d255-seed1-demo", 1242, 375, 255, 1, lambda d: checkers.demo(d)),
    # parameter corners of the fused kernels (same cases as tests/test_oracle.py pins against the reference)
    ("gap0-nomean", 320, 160, 63, 7, lambda d: checkers.stereomapper(d).copy(ipol_gap_width=0, filter_adaptive_mean=0)),
    ("gap2-both", 320, 160, 63, 8, lambda d: checkers.demo(d).copy(ipol_gap_width=2)),
    ("radius3", 320, 160, 63, 9, lambda d: checkers.stereomapper(d).copy(sradius=3.0, match_texture=40)),
    ("grid16", 320, 160, 63, 10, lambda d: checkers.stereomapper(d).copy(grid_size=16, speckle_size=50)),
    ("step4-duplicate-vertices", 320, 160, 63, 11, lambda d: checkers.stereomapper(d).copy(candidate_stepsize=4)),
]


@pytest.mark.parametrize("tag,W,H,dmax,seed,mk", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle_all_stages(oracle, tag, W, H, dmax, seed, mk):
    L, R, _ = synth.synthetic_pair(W, H, dmax, seed)
    p = mk(dmax)
    rc_o, _, _, st_o = oracle.run_stages(L, R, p)
    rc, D1, D2, st = run_cuda(L, R, p)
    assert rc == rc_o == 0
    check_against(st, st_o, D1, D2, tag)


@pytest.mark.parametrize("mesh", ["device", "host"])
def test_hd_config_all_stages(oracle, monkeypatch, mesh):
    """BASELINE config 2: 1920x1080, d_max 128, with the mesh stage (lattice filters + Delaunay) on the GPU -- the
    triangulation of ~4000 points works out of the global scratch area -- and on the host."""
    monkeypatch.setenv("ELAS_B200_HOST_STAGE", "0" if mesh == "device" else "1")
    L, R, _ = synth.synthetic_pair(1920, 1080, 128, 0)
    p = checkers.stereomapper(128)
    rc_o, _, _, st_o = oracle.run_stages(L, R, p)
    rc, D1, D2, st = run_cuda(L, R, p)
    assert rc == rc_o == 0
    check_against(st, st_o, D1, D2, "HD")


def test_drop_in_call_matches_oracle(oracle):
    """elas_b200_process: the synchronous entry a replacement Elas::process binds."""
    L, R, _ = synth.synthetic_pair(416, 200, 95, 4)
    p = checkers.stereomapper(95)
    _, O1, O2 = oracle.process(L, R, p)
    rc, D1, D2 = elas_b200.process(L, R, as_product_params(p))
    assert rc == 0 and bits_equal(D1, O1) and bits_equal(D2, O2)
    # again through the cached context, with a wider row stride (stereothread.cpp:111 passes widthStep)
    Lp = np.zeros((200, 420), np.uint8); Lp[:, :416] = L
    Rp = np.zeros((200, 420), np.uint8); Rp[:, :416] = R
    rc, E1, E2 = elas_b200.process(Lp[:, :416], Rp[:, :416], as_product_params(p))
    assert rc == 0 and bits_equal(E1, O1) and bits_equal(E2, O2)


def test_too_few_support_points():
    """elas.cpp:69-75: the reference returns without writing D; the C ABI returns 1 and fills -10."""
    flat = np.full((100, 160), 90, np.uint8)
    rc, D1, D2 = elas_b200.process(flat, flat, elas_b200.stereomapper(63))
    assert rc == 1 and (D1 == -10).all() and (D2 == -10).all()


def test_batch_pipeline_equals_single_frames(oracle):
    """n frames over 4 slots (the throughput path) give exactly the per-frame results, in order."""
    p = checkers.stereomapper(95)
    pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in range(10, 19)]
    e = elas_b200.ElasB200(as_product_params(p), 416, 200, n_slots=4)
    try:
        status, D1, D2 = e.process_batch([a for a, _ in pairs], [b for _, b in pairs])
        launches = e.launch_count()
    finally:
        e.close()
    assert status == [0] * len(pairs) and launches >= 9 * len(pairs)      # 9 kernels per frame
    for i, (L, R) in enumerate(pairs):
        _, O1, O2 = oracle.process(L, R, p)
        assert bits_equal(D1[i], O1) and bits_equal(D2[i], O2), f"frame {i}"


def test_idempotent_and_slot_independent():
    """Size-independent property at the full K size: same pair through two slots, twice -> same bits."""
    L, R, _ = synth.synthetic_pair(1242, 375, 255, 3)
    e = elas_b200.ElasB200(elas_b200.stereomapper(255), 1242, 375, n_slots=2)
    try:
        _, A1, A2 = e.process(L, R, slot=0)
        _, B1, B2 = e.process(L, R, slot=1)
        _, C1, C2 = e.process(L, R, slot=0)
    finally:
        e.close()
    assert bits_equal(A1, B1) and bits_equal(A2, B2) and bits_equal(A1, C1) and bits_equal(A2, C2)
    valid = A1 >= 0
    assert 0.5 < valid.mean() < 0.95
    # disparities are bounded by the parameter block; invalid pixels carry exactly -10
    assert A1.max() <= 255 and (A1[~valid] == -10).all() and (A2[A2 < 0] == -10).all()


def test_left_right_symmetry_property():
    """Mirroring both images and swapping them swaps the roles of D1 and D2 (mirrored)."""
    L, R, _ = synth.synthetic_pair(640, 240, 127, 5)
    p = elas_b200.demo(127)
    _, D1, D2 = elas_b200.process(L, R, p)
    _, M1, M2 = elas_b200.process(np.ascontiguousarray(R[:, ::-1]), np.ascontiguousarray(L[:, ::-1]), p)
    # the candidate lattice is anchored at u=0, so the mirrored problem is not bit-identical; the
    # valid overlap must agree within one level almost everywhere
    a, b = D2, M1[:, ::-1]
    both = (a >= 0) & (b >= 0)
    assert both.mean() > 0.4 and (np.abs(a[both] - b[both]) <= 1).mean() > 0.97


def test_reference_available_on_box_matches(ref):
    """When the prebuilt reference travelled with the snapshot, check the CUDA path against it too."""
    L, R, _ = synth.synthetic_pair(1242, 375, 255, 2)
    p = checkers.stereomapper(255)
    _, R1, R2 = ref.process(L, R, p)
    rc, D1, D2 = elas_b200.process(L, R, as_product_params(p))
    assert rc == 0 and bits_equal(D1, R1) and bits_equal(D2, R2)


def test_cpp_drop_in_class_matches_oracle(oracle, tmp_path):
    """The C++ drop-in `Elas` class (stereo-vision_b200/dropin/libelas/src), compiled into a headless
    stand-in for the reference's call sites (stereothread.cpp:76-114, main.cpp:61-64), gives the
    oracle's maps bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "stereo-vision_b200", "dropin", "libelas", "src")
    exe = str(tmp_path / "dropin_demo")
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-I" + src, os.path.join(root, "tests", "dropin_demo.cpp"),
                           os.path.join(src, "elas.cpp"), os.path.join(src, "descriptor.cpp"), "-ldl", "-o", exe])
    W, H, dmax = 640, 240, 127
    L, R, _ = synth.synthetic_pair(W, H, dmax, 21)
    L.tofile(str(tmp_path / "l.raw")); R.tofile(str(tmp_path / "r.raw"))
    env = dict(os.environ, ELAS_B200_LIB=elas_b200.LIB_PATH)
    for mode, p in (("stereomapper", checkers.stereomapper(dmax)), ("demo", checkers.demo(dmax))):
        subprocess.check_call([exe, mode, str(W), str(H), str(dmax), str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                               str(tmp_path / "d1.out"), str(tmp_path / "d2.out")], env=env)
        D1 = np.fromfile(str(tmp_path / "d1.out"), np.float32).reshape(H, W)
        D2 = np.fromfile(str(tmp_path / "d2.out"), np.float32).reshape(H, W)
        _, O1, O2 = oracle.process(L, R, p)
        assert bits_equal(D1, O1) and bits_equal(D2, O2), mode


def test_batch_scheduler_more_slots_than_workers_and_a_blank_frame(oracle):
    """The batch scheduler (workers decoupled from slots): 20 frames over 6 slots driven by 2 workers,
    one frame without texture (fewer than 3 support points: status 1, maps all -10).  Results must be
    the per-frame oracle results, in order, whatever slot and worker handled a frame."""
    p = checkers.stereomapper(95)
    pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in (21, 22, 23)]
    blank = (np.full((200, 416), 90, np.uint8), np.full((200, 416), 90, np.uint8))
    order = [0, 1, 2, 0, 3, 1, 2, 2, 0, 1, 3, 0, 1, 2, 0, 1, 2, 0, 1, 2]
    frames = [pairs[i] if i < 3 else blank for i in order]
    e = elas_b200.ElasB200(as_product_params(p), 416, 200, n_slots=6, n_workers=2)
    try:
        status, D1, D2 = e.process_batch([a for a, _ in frames], [b for _, b in frames])
        status2, E1, E2 = e.process_batch([a for a, _ in frames], [b for _, b in frames])   # context reuse
    finally:
        e.close()
    want = [oracle.process(L, R, p) for L, R in pairs]
    for i, k in enumerate(order):
        if k == 3:
            assert status[i] == elas_b200.E_FEW_SUPPORT and (D1[i] == -10).all() and (D2[i] == -10).all()
        else:
            assert status[i] == 0 and bits_equal(D1[i], want[k][1]) and bits_equal(D2[i], want[k][2]), f"frame {i}"
        assert status2[i] == status[i] and bits_equal(E1[i], D1[i]) and bits_equal(E2[i], D2[i])


def test_bandwidth_config_4096x2160_all_stages(oracle, monkeypatch):
    """BASELINE.json configs[4] geometry at full size (the oracle takes ~10 s): every stage against the
    oracle with the mesh stage on the GPU (820x432 lattice and ~14 600 support points worked on in global
    memory), then the size-independent properties with the host stage -- two slots agree bit for bit, the batch
    path agrees with the single-frame path, D2 (L/R-checked only) holds integers."""
    W, H, dmax = 4096, 2160, 256
    L, R, _ = synth.synthetic_pair(W, H, dmax, 1)
    p = checkers.stereomapper(dmax)
    rc_o, _, _, st_o = oracle.run_stages(L, R, p)
    monkeypatch.setenv("ELAS_B200_HOST_STAGE", "0")
    rc, A1, A2, st = run_cuda(L, R, p)
    assert rc == rc_o == 0
    check_against(st, st_o, A1, A2, "4096x2160")
    del st, st_o
    monkeypatch.setenv("ELAS_B200_HOST_STAGE", "1")
    e = elas_b200.ElasB200(elas_b200.stereomapper(dmax), W, H, n_slots=2, n_workers=2)
    try:
        rc, B1, B2 = e.process(L, R, slot=1)
        status, C1, C2 = e.process_batch([L, L], [R, R])
    finally:
        e.close()
    assert rc == 0 and status == [0, 0]
    assert bits_equal(A1, B1) and bits_equal(A2, B2)
    for k in range(2):
        assert bits_equal(C1[k], A1) and bits_equal(C2[k], A2)
    v1, v2 = A1 >= 0, A2 >= 0
    assert 0.8 < v1.mean() < 0.99 and 0.8 < v2.mean() < 0.99
    assert A1.max() <= dmax and A2.max() <= dmax and (A1[~v1] == -10).all() and (A2[~v2] == -10).all()
    assert (A2[v2] == np.round(A2[v2])).all()


def test_device_pointers_through_the_host_buffer_entry_point():
    """elas_b200_process_batch (the host-buffer entry) must also accept device pointers for any of the four
    buffers (cudaMemcpyDefault semantics): the int16 transfer + CPU widening of D2 applies to host memory only."""
    torch = pytest.importorskip("torch")
    W, H, dmax = 416, 200, 95
    L, R, _ = synth.synthetic_pair(W, H, dmax, 31)
    p = elas_b200.stereomapper(dmax)
    dI = torch.stack([torch.from_numpy(L), torch.from_numpy(R)]).cuda()
    dD = torch.full((2, 2, H, W), -77.0, dtype=torch.float32, device="cuda")
    hD = torch.full((2, H, W), -77.0, dtype=torch.float32).pin_memory()
    e = elas_b200.ElasB200(p, W, H, n_slots=2, n_workers=1)
    try:
        rc, D1, D2 = e.process(L, R)
        # frame 0: device in, device out; frame 1: device in, pinned host out
        st = e.process_batch_ptrs([dI[0].data_ptr(), dI[0].data_ptr()], [dI[1].data_ptr(), dI[1].data_ptr()],
                                  [dD[0, 0].data_ptr(), hD[0].data_ptr()], [dD[0, 1].data_ptr(), hD[1].data_ptr()], W, device=False)
        torch.cuda.synchronize()
    finally:
        e.close()
    assert rc == 0 and st == [0, 0]
    assert bits_equal(dD[0, 0].cpu().numpy(), D1) and bits_equal(dD[0, 1].cpu().numpy(), D2)
    assert bits_equal(hD[0].numpy(), D1) and bits_equal(hD[1].numpy(), D2)


def test_frame_groups_equal_single_frames(oracle):
    """Frames batched per launch chain (every kernel takes the frame as a grid dimension): 19 frames over 3 groups
    of 4 frames driven by 2 workers, one frame without texture; results are the per-frame oracle results, in order."""
    p = checkers.stereomapper(95)
    pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in (41, 42, 43, 44, 45)]
    blank = (np.full((200, 416), 90, np.uint8), np.full((200, 416), 90, np.uint8))
    order = [0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 1, 0, 5, 2, 3, 4, 0, 1]
    frames = [pairs[i] if i < 5 else blank for i in order]
    e = elas_b200.ElasB200(as_product_params(p), 416, 200, n_slots=3, n_workers=2, frames_per_group=4)
    try:
        assert e.frames_per_group == 4 and e.mesh_on_device
        status, D1, D2 = e.process_batch([a for a, _ in frames], [b for _, b in frames])
        status2, E1, E2 = e.process_batch([a for a, _ in frames[:5]], [b for _, b in frames[:5]])   # a partial last group
    finally:
        e.close()
    want = [oracle.process(L, R, p) for L, R in pairs]
    for i, k in enumerate(order):
        if k == 5:
            assert status[i] == elas_b200.E_FEW_SUPPORT and (D1[i] == -10).all() and (D2[i] == -10).all()
        else:
            assert status[i] == 0 and bits_equal(D1[i], want[k][1]) and bits_equal(D2[i], want[k][2]), f"frame {i}"
    for i in range(5):
        assert status2[i] == 0 and bits_equal(E1[i], want[i][1]) and bits_equal(E2[i], want[i][2])


def test_frame_groups_at_the_metric_size_device_buffers(oracle):
    """1242x375 d_max 255, groups of 8 frames, device-resident inputs and outputs (the bench's `value` path)."""
    torch = pytest.importorskip("torch")
    W, H, dmax = 1242, 375, 255
    p = checkers.stereomapper(dmax)
    pairs = [synth.synthetic_pair(W, H, dmax, s)[:2] for s in (50, 51, 52)]
    n = 20
    dI = torch.stack([torch.stack([torch.from_numpy(pairs[i % 3][0]), torch.from_numpy(pairs[i % 3][1])]) for i in range(n)]).cuda()
    dD = torch.full((n, 2, H, W), -77.0, dtype=torch.float32, device="cuda")
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=2, n_workers=1, frames_per_group=0)
    try:
        assert e.frames_per_group == 8
        st = e.process_batch_ptrs([dI[i, 0].data_ptr() for i in range(n)], [dI[i, 1].data_ptr() for i in range(n)],
                                  [dD[i, 0].data_ptr() for i in range(n)], [dD[i, 1].data_ptr() for i in range(n)], W, device=True)
        torch.cuda.synchronize()
    finally:
        e.close()
    assert st == [0] * n
    out = dD.cpu().numpy()
    want = [oracle.process(L, R, p) for L, R in pairs]
    for i in range(n):
        assert bits_equal(out[i, 0], want[i % 3][1]) and bits_equal(out[i, 1], want[i % 3][2]), f"frame {i}"


def test_host_stage_path_equals_device_mesh_stage(oracle, monkeypatch):
    """Contexts whose parameters the device mesh stage does not take run lattice filters and Delaunay on the host
    (host_stage.cc); forcing that path for a ROBOTICS-family set gives the same bits."""
    L, R, _ = synth.synthetic_pair(640, 240, 127, 33)
    p = checkers.stereomapper(127)
    _, O1, O2 = oracle.process(L, R, p)
    monkeypatch.setenv("ELAS_B200_HOST_STAGE", "1")
    e = elas_b200.ElasB200(as_product_params(p), 640, 240, n_slots=2, n_workers=2, frames_per_group=2)
    try:
        assert not e.mesh_on_device
        status, D1, D2 = e.process_batch([L] * 5, [R] * 5)
    finally:
        e.close()
    assert status == [0] * 5
    for i in range(5):
        assert bits_equal(D1[i], O1) and bits_equal(D2[i], O2)
    # MIDDLEBURY adds corner support points (elas.cpp:283-318): host stage by construction
    e = elas_b200.ElasB200(elas_b200.middlebury().copy(disp_max=63), 320, 160, n_slots=1)
    try:
        assert not e.mesh_on_device
    finally:
        e.close()


def test_multi_device_entry_shards_frames(oracle):
    """elas_b200_multi_*: one process, one context per device (here the same device twice when the box has one GPU),
    frame i -> context i mod n; results equal the per-frame oracle results, in order."""
    n_dev = elas_b200.load_library().elas_b200_device_count()
    devices = [0, 1] if n_dev >= 2 else [0, 0]
    p = checkers.stereomapper(95)
    pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in (61, 62, 63)]
    frames = [pairs[i % 3] for i in range(11)]
    m = elas_b200.ElasB200Multi(as_product_params(p), 416, 200, devices, n_groups=2, frames_per_group=2, n_workers=1)
    try:
        status, D1, D2 = m.process_batch([a for a, _ in frames], [b for _, b in frames])
    finally:
        m.close()
    want = [oracle.process(L, R, p) for L, R in pairs]
    assert status == [0] * 11
    for i in range(11):
        assert bits_equal(D1[i], want[i % 3][1]) and bits_equal(D2[i], want[i % 3][2]), f"frame {i}"


def test_drop_in_call_cache_evicts_and_recreates(oracle):
    """elas_b200_process keeps at most four cached contexts (one per device / size / parameter block); a seventh
    combination evicts the least recently used one, and coming back to an evicted combination builds it again."""
    sizes = [(160 + 16 * k, 96 + 8 * k) for k in range(6)]
    first = None
    for rnd in range(2):
        for W, H in sizes:
            L, R, _ = synth.synthetic_pair(W, H, 31, seed=W)
            rc, D1, D2 = elas_b200.process(L, R, elas_b200.stereomapper(31))
            assert rc == 0
            if (W, H) == sizes[0]:
                if first is None:
                    _, O1, O2 = oracle.process(L, R, checkers.stereomapper(31))
                    assert np.array_equal(D1.view(np.uint32), O1.view(np.uint32)) and np.array_equal(D2.view(np.uint32), O2.view(np.uint32))
                    first = D1
                else:
                    assert np.array_equal(D1.view(np.uint32), first.view(np.uint32))


def test_frame_groups_at_hd(oracle):
    """1920x1080 d_max 128: four frames per launch chain (the default group size at this frame size), the triangulations
    of a chain working side by side in the global scratch area; six frames = one full and one partial chain."""
    W, H, dmax = 1920, 1080, 128
    p = checkers.stereomapper(dmax)
    pairs = [synth.synthetic_pair(W, H, dmax, s)[:2] for s in (60, 61)]
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=1, n_workers=1, frames_per_group=0)
    try:
        assert e.frames_per_group == 4 and e.mesh_on_device
        status, D1, D2 = e.process_batch([pairs[i % 2][0] for i in range(6)], [pairs[i % 2][1] for i in range(6)])
    finally:
        e.close()
    want = [oracle.process(L, R, p) for L, R in pairs]
    for i in range(6):
        assert status[i] == 0 and bits_equal(D1[i], want[i % 2][1]) and bits_equal(D2[i], want[i % 2][2]), f"frame {i}"


def test_full_chain_after_a_short_chain(oracle):
    """One group of four frames: chains of 4, 3 and 4 frames.  A chain leaves the grid scatter planes of its own
    frames zeroed for the chain after it; frame 3 of the third chain must not see what the first chain left there."""
    p = checkers.stereomapper(95)
    pairs = [synth.synthetic_pair(416, 200, 95, s)[:2] for s in (71, 72, 73, 74)]
    want = [oracle.process(L, R, p) for L, R in pairs]
    e = elas_b200.ElasB200(as_product_params(p), 416, 200, n_slots=1, n_workers=1, frames_per_group=4)
    try:
        for n in (4, 3, 4, 1, 2, 4):
            status, D1, D2 = e.process_batch([pairs[i][0] for i in range(n)], [pairs[i][1] for i in range(n)])
            for i in range(n):
                assert status[i] == 0 and bits_equal(D1[i], want[i][1]) and bits_equal(D2[i], want[i][2]), f"chain of {n}, frame {i}"
    finally:
        e.close()


def _shifted_pair(W, H, d_const, seed):
    """A textured pair whose lower half has the constant disparity d_const (up to disp_max itself) and whose upper
    half has disparity 7: left(u) = texture(u - d)."""
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (H, W + d_const), dtype=np.uint8)
    tex = (0.5 * tex + 0.5 * np.repeat(np.repeat(rng.integers(0, 256, (H // 4 + 1, (W + d_const) // 4 + 1)), 4, 0), 4, 1)[:H, :W + d_const]).astype(np.uint8)
    R = tex[:, d_const:d_const + W].copy()
    L = np.empty_like(R)
    L[: H // 2] = tex[: H // 2, d_const - 7:d_const - 7 + W]
    L[H // 2:] = tex[H // 2:, :W]
    return L, R


@pytest.mark.parametrize("mode", ["2", "1", "0"])
def test_narrowed_right_map_reaches_disp_max(oracle, monkeypatch, mode):
    """The right map crosses PCIe as u8 + validity bits (ELAS_B200_NARROW_D2=2, disp_max <= 255), as int16 (=1) or as
    float32 (=0); a scene at disparity 255 = disp_max puts the largest byte value next to invalid pixels.  Unaligned
    caller buffers and a width that is no multiple of 32 exercise the edges of the widening code."""
    monkeypatch.setenv("ELAS_B200_NARROW_D2", mode)
    W, H, dmax = 629, 121, 255      # an odd pixel count: the frames of a group sit at odd multiples of 2 bytes
    L, R = _shifted_pair(W, H, 255, 5)
    p = checkers.stereomapper(dmax)
    _, O1, O2 = oracle.process(L, R, p)
    assert (O2 == 255).sum() > 1000 and (O2 == -10).sum() > 1000
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=2, n_workers=1, frames_per_group=2)
    try:
        # caller's maps at addresses that are 4 mod 16
        store = [np.zeros(W * H + 8, np.float32) for _ in range(6)]
        maps = [a[(1 + (-a.ctypes.data // 4)) % 4:][:W * H].reshape(H, W) for a in store]
        assert all(m.ctypes.data % 16 == 4 for m in maps)
        st = e.process_batch_ptrs([L.ctypes.data] * 3, [R.ctypes.data] * 3, [m.ctypes.data for m in maps[:3]],
                                  [m.ctypes.data for m in maps[3:]], L.strides[0])
    finally:
        e.close()
    assert st == [0, 0, 0]
    for i in range(3):
        assert bits_equal(maps[i], O1) and bits_equal(maps[3 + i], O2), f"mode {mode} frame {i}"


def test_narrowed_right_map_beyond_a_byte(oracle):
    """disp_max = 300: the narrowed right map falls back to int16."""
    W, H, dmax = 700, 96, 300
    L, R = _shifted_pair(W, H, 290, 6)
    p = checkers.stereomapper(dmax)
    _, O1, O2 = oracle.process(L, R, p)
    assert (O2 > 255).sum() > 500
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=1, n_workers=1)
    try:
        st, D1, D2 = e.process_batch([L, L], [R, R])
    finally:
        e.close()
    assert st == [0, 0] and bits_equal(D1[1], O1) and bits_equal(D2[1], O2)


def test_device_buffers_through_the_unfused_tail(oracle):
    """Device-resident maps with a parameter set that runs the unfused post-processing chain (median): the maps are
    reached by device-to-device copies -- in particular the right map must not take the narrowed host path, whose
    CPU-side widening cannot write device memory."""
    torch = pytest.importorskip("torch")
    W, H, dmax = 512, 188, 100
    p = checkers.stereomapper(dmax).copy(filter_median=1)
    L, R, _ = synth.synthetic_pair(W, H, dmax, 81)
    _, O1, O2 = oracle.process(L, R, p)
    n = 5
    dI = torch.stack([torch.stack([torch.from_numpy(L), torch.from_numpy(R)]) for _ in range(n)]).cuda()
    dD = torch.full((n, 2, H, W), -77.0, dtype=torch.float32, device="cuda")
    e = elas_b200.ElasB200(as_product_params(p), W, H, n_slots=2, n_workers=2, frames_per_group=3)
    try:
        st = e.process_batch_ptrs([dI[i, 0].data_ptr() for i in range(n)], [dI[i, 1].data_ptr() for i in range(n)],
                                  [dD[i, 0].data_ptr() for i in range(n)], [dD[i, 1].data_ptr() for i in range(n)], W, device=True)
        torch.cuda.synchronize()
        # the same buffers through the HOST entry point (pointer classification instead of the device flag)
        dD2 = torch.full((n, 2, H, W), -77.0, dtype=torch.float32, device="cuda")
        st2 = e.process_batch_ptrs([dI[i, 0].data_ptr() for i in range(n)], [dI[i, 1].data_ptr() for i in range(n)],
                                   [dD2[i, 0].data_ptr() for i in range(n)], [dD2[i, 1].data_ptr() for i in range(n)], W, device=False)
        torch.cuda.synchronize()
    finally:
        e.close()
    assert st == [0] * n and st2 == [0] * n
    for out in (dD.cpu().numpy(), dD2.cpu().numpy()):
        for i in range(n):
            assert bits_equal(out[i, 0], O1) and bits_equal(out[i, 1], O2), f"frame {i}"


def test_sixteen_frames_per_chain(oracle):
    """The largest frame group (opt-in, frames_per_group = 16): 21 frames = one full chain and a tail of 5."""
    p = checkers.stereomapper(63)
    pairs = [synth.synthetic_pair(320, 120, 63, s)[:2] for s in (91, 92, 93)]
    want = [oracle.process(L, R, p) for L, R in pairs]
    e = elas_b200.ElasB200(as_product_params(p), 320, 120, n_slots=1, n_workers=1, frames_per_group=16)
    try:
        assert e.frames_per_group == 16
        for n in (21, 16, 3):
            status, D1, D2 = e.process_batch([pairs[i % 3][0] for i in range(n)], [pairs[i % 3][1] for i in range(n)])
            for i in range(n):
                assert status[i] == 0 and bits_equal(D1[i], want[i % 3][1]) and bits_equal(D2[i], want[i % 3][2]), f"batch of {n}, frame {i}"
    finally:
        e.close()
