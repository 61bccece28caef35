#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the dense-stereo hot path (Elas::process) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one batch of B synthetic 1242x375 stereo pairs (d_max 255, stereomapper's parameter
set: BASELINE.json configs[1]) per GPU through the whole path.  Frames are independent, so they are
sharded over ranks with no data-path collective (weak scaling: B pairs per GPU per step); the only
collective is one NCCL broadcast of the parameter block at start-up.

Printed JSON (rank 0, one line):
  value      pairs/s, whole job, inputs/outputs resident in HBM (elas_b200_process_batch_device)
  e2e        pairs/s through the C ABI with PINNED HOST buffers (elas_b200_process_batch): the
             host->device copies of both images and device->host copies of both disparity maps are
             inside the timed region
  roofline   matching kernel (K7, elas.cpp:960-1118): algorithmic bytes / CUDA-event time per launch
             against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference libelas (oracle/_ref, -O3 -msse3) on the host cores of this
             box, one process per core, bounded sample
--impl reference times only that CPU arm and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
# 32 hardware work queues instead of 8 (read at CUDA context creation, i.e. before torch touches the device): the frame
# groups' launch chains run on separate streams and must not queue behind each other's long single-CTA kernels
# (libelas_b200.so sets the same default when it is loaded first, elas_b200.cu)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

# BASELINE.json configs: K = configs[1] (the metric's configuration, default), HD = configs[2], 4K = configs[4] geometry
CONFIGS = {"K": (1242, 375, 255, 512, 8), "HD": (1920, 1080, 128, 256, 2), "4K": (4096, 2160, 256, 64, 1)}
W, H, DMAX = CONFIGS["K"][:3]
WORKLOAD = f"synthetic {W}x{H} random-texture stereo pairs, d_max={DMAX}, stereomapper parameter set"
METRIC = "stereo pairs/sec @1242x375 d_max=255"


def select_config(name):
    global W, H, DMAX, WORKLOAD, METRIC
    W, H, DMAX = CONFIGS[name][:3]
    WORKLOAD = f"synthetic {W}x{H} random-texture stereo pairs, d_max={DMAX}, stereomapper parameter set"
    METRIC = f"stereo pairs/sec @{W}x{H} d_max={DMAX}"


def algorithmic_bytes_matching(w, h, dmax, grid_size=20):
    """B_match = 72*N + 8*gw*gh*(dmax+2) bytes per stereo pair (both directions)."""
    gw, gh = -(-w // grid_size), -(-h // grid_size)
    return 72 * w * h + 8 * gw * gh * (dmax + 2)


def ncu_traffic_k7(which=None, frames_per_launch=None):
    """DRAM bytes per launch of the matching kernel from the committed ncu --set full captures
    (which = None: the 1242x375 workload; "bandwidth_config": 4096x2160), scaled to the frames a launch of this run
    processes (the captures are per launch of `frames_per_launch` frames)."""
    path = os.path.join(ROOT, "profiles", "r02_k7_traffic.json")
    if not os.path.exists(path):
        return None
    rec = json.load(open(path))
    if which:
        rec = rec.get(which)
        if not rec:
            return None
    total = int(rec["dram_bytes_read_per_launch"]) + int(rec["dram_bytes_write_per_launch"])
    captured = int(rec.get("frames_per_launch", 1))
    return total if not frames_per_launch else int(round(total * frames_per_launch / captured))


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md): nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref), one PROCESS per core (Triangle keeps file-scope
# state, triangle.cpp:541-550, so threads would race)
# ------------------------------------------------------------------------------------------------
_cpu_state = {}


def _cpu_init(kind, config):
    import checkers
    select_config(config)
    _cpu_state["impl"] = checkers.RefElas() if kind == "reference" else checkers.OracleElas()
    _cpu_state["params"] = checkers.stereomapper(DMAX)


def _cpu_work(args):
    seed, n = args
    import synth
    L, R, _ = synth.synthetic_pair(W, H, DMAX, seed)
    impl, p = _cpu_state["impl"], _cpu_state["params"]
    t0 = time.perf_counter()
    for _ in range(n):
        rc, D1, D2 = impl.process(L, R, p)
    return time.perf_counter() - t0, int((D1 >= 0).sum())


def cpu_kind():
    import checkers
    return "reference" if checkers.have_ref() else "port"


class CpuArm:
    """A pool of one worker process per host core, each with the checker library loaded."""

    def __init__(self, config="K"):
        self.kind = cpu_kind()
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(self.kind, config))
        self.pool.map(_cpu_work, [(0, 1)] * self.cores)      # load libraries, touch memory

    def run(self, pairs_per_core):
        t0 = time.perf_counter()
        self.pool.map(_cpu_work, [(s, pairs_per_core) for s in range(self.cores)], chunksize=1)
        dt = time.perf_counter() - t0
        return self.cores * pairs_per_core / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, rank):
    if rank != 0:
        return
    arm = CpuArm(args.config)
    per_core = args.cpu_pairs_per_core or CONFIGS[args.config][4]
    for _ in range(args.warmup):
        arm.run(1)
    t_total, pairs = 0.0, 0
    for _ in range(args.steps):
        rate, dt = arm.run(per_core)
        t_total += dt
        pairs += arm.cores * per_core
    arm.close()
    value = pairs / t_total
    sample = (f"{arm.cores} processes x {per_core} pairs per step, {args.steps} steps, Elas::process only "
              "(its _mm_malloc buffers come zero-filled from the harness, a few % of extra memset, oracle/ref_harness.cpp)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": arm.cores * per_core, "host_cores": arm.cores},
        "cpu_baseline": {"value": round(value, 3), "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="K", choices=sorted(CONFIGS), help="K = 1242x375 d255 (the metric), HD = 1920x1080 d128, 4K = 4096x2160 d256")
    ap.add_argument("--batch", type=int, default=0, help="stereo pairs per GPU per step (0 = the configuration's default)")
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic pairs cycled through the batch")
    ap.add_argument("--slots", type=int, default=0, help="frame groups in flight per GPU (0 = three per worker)")
    ap.add_argument("--group", type=int, default=0, help="frames per launch chain (0 = chosen from the frame size)")
    ap.add_argument("--workers", type=int, default=0, help="host worker threads per GPU (0 = from the cores per rank)")
    ap.add_argument("--cpu-pairs-per-core", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-4k", action="store_true", help="skip the extra roofline points of the matching kernel (HD, 4096x2160)")
    args = ap.parse_args()
    select_config(args.config)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import elas_b200
    import sharding
    import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # the one collective of the path: rank 0's parameter block to every rank (NCCL broadcast)
    params = sharding.broadcast_params(elas_b200.stereomapper(DMAX), dev, src=0)

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # host workers: this rank's share of the cores; slots: twice that, so the GPU has phases queued while
    # every worker runs a host stage (workers are not tied to slots, elas_b200.cu)
    share = cores // max(world, 1)
    # The whole frame runs on the GPU, so a worker only enqueues launch chains and collects results (the end-to-end
    # path also widens D2 from int16 on the host): a few per GPU, one core per rank stays free for the main thread
    workers = args.workers or max(1, min(8, share - 1))
    slots = args.slots or max(2, min(32, 4 * workers))
    B = args.batch or CONFIGS[args.config][3]
    args.distinct = min(args.distinct, B)
    bpl = W + 15 - (W - 1) % 16

    # synthetic inputs: `distinct` seeded pairs per rank, cycled to fill the batch
    pairs = [synth.synthetic_pair(W, H, DMAX, seed=1000 * rank + i)[:2] for i in range(args.distinct)]
    # pinned host buffers of the end-to-end path; if the box refuses to pin that much, halve the batch
    # (every rank must take the same decision: weak scaling keeps B equal across ranks)
    while True:
        try:
            h_I = torch.zeros((B, 2, H, bpl), dtype=torch.uint8).pin_memory()
            h_D = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
            ok = 1
        except RuntimeError:
            h_I = h_D = None
            ok = 0
        (all_ok,) = sharding.sum_over_ranks([ok], dev)
        if int(all_ok) == world:
            break
        h_I = h_D = None
        B //= 2
        if B < 16:
            raise SystemExit("bench.py: cannot pin the host buffers of the end-to-end path")
    for i in range(B):
        L, R = pairs[i % args.distinct]
        h_I[i, 0, :, :W] = torch.from_numpy(L)
        h_I[i, 1, :, :W] = torch.from_numpy(R)
    d_I = h_I.to(dev)
    d_D = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # 4x the 126 MB L2

    def ptrs(t, k):
        return [t[i, k].data_ptr() for i in range(B)]

    engine = elas_b200.ElasB200(params, W, H, n_slots=slots, device=local_rank, n_workers=workers, frames_per_group=args.group)
    if not engine.mesh_on_device:
        # lattice filters + Delaunay on the host for this geometry (the lattice does not fit one CTA's shared memory):
        # one worker per core pays off again
        engine.close()
        workers = args.workers or max(1, share - 1)
        slots = args.slots or max(2, min(32, 2 * workers))
        engine = elas_b200.ElasB200(params, W, H, n_slots=slots, device=local_rank, n_workers=workers, frames_per_group=args.group)

    # pointer tables of the batch, built once: the timed call is the C ABI call and nothing else
    dev_ptrs = (ptrs(d_I, 0), ptrs(d_I, 1), ptrs(d_D, 0), ptrs(d_D, 1))
    host_ptrs = (ptrs(h_I, 0), ptrs(h_I, 1), ptrs(h_D, 0), ptrs(h_D, 1))

    def step_device():
        return engine.process_batch_ptrs(*dev_ptrs, bpl, device=True)

    def step_host():
        return engine.process_batch_ptrs(*host_ptrs, bpl, device=False)

    def step_host_left_only():
        # opt-in: the right map is not returned (D2[i] == NULL); stereomapper reads only D1 (stereothread.cpp:116-147)
        return engine.process_batch_ptrs(host_ptrs[0], host_ptrs[1], host_ptrs[2], [0] * B, bpl, device=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        """K steps, each bracketed by CUDA events on the current stream (which is idle, so an event
        completes when recorded and the pair spans all slot streams of the step); L2 flushed between."""
        total_ms, bad = 0.0, 0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            status = step_fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            bad += sum(1 for s in status if s != 0)
        return total_ms, bad

    # warm-up (both paths), then the two timed regions
    for _ in range(max(args.warmup, 3)):
        step_device()
        step_host()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = engine.launch_count()
    ms_dev, bad_dev = timed(step_device, args.steps)
    launches = engine.launch_count() - launches0
    barrier()
    ms_host, bad_host = timed(step_host, args.steps)
    barrier()
    ms_left, bad_left = timed(step_host_left_only, max(2, args.steps // 2))
    ms_left /= max(2, args.steps // 2)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # parity guards inside the bench: device-resident and host paths must give identical maps, and every rank
    # checks frames of its own against the CPU oracle (bit for bit)
    same = bool(torch.equal(d_D.cpu(), h_D))
    import checkers
    oracle = checkers.OracleElas()
    oracle_mismatch, oracle_checked = 0, 0
    for i in sorted({0, B // 2, B - 1} if W * H < 4000000 else {0}):
        L, R = pairs[i % args.distinct]
        _, O1, O2 = oracle.process(L, R, checkers.Params.from_buffer_copy(bytes(params)))
        got = h_D[i].numpy()
        oracle_checked += 1
        if not (np.array_equal(got[0].view(np.uint32), O1.view(np.uint32)) and np.array_equal(got[1].view(np.uint32), O2.view(np.uint32))):
            oracle_mismatch += 1
    oracle_mismatch, oracle_checked = (int(x) for x in sharding.sum_over_ranks([oracle_mismatch, oracle_checked], dev))

    ms_dev_max, ms_host_max, ms_left_max = sharding.max_over_ranks([ms_dev, ms_host, ms_left], dev)
    narrow_mode = int(os.environ.get("ELAS_B200_NARROW_D2", "2"))
    d2_bytes = 4 * W * H if narrow_mode == 0 else 2 * W * H if (narrow_mode == 1 or DMAX > 255) else \
        ((W * H + 15) // 16 * 16 + H * ((W + 31) // 32) * 4)
    d2h_bytes_per_pair = 4 * W * H + d2_bytes

    # roofline of the matching kernel: isolated launches cycling over all slots' tables (their
    # combined descriptors exceed L2), CUDA events on the launching stream, L2 flushed first
    k7_ms, k7_frames = engine.time_matching(iters=30, flush_l2=True, per_frame=False)
    peak, peak_src = measured_hbm_peak()
    b_match = algorithmic_bytes_matching(W, H, DMAX) * k7_frames          # one launch processes a frame group
    achieved = b_match / (k7_ms * 1e-3) / 1e9

    # the consumers of D1 (colour map, back-projection) on a frame left in HBM:
    # algorithmic bytes 16 N (4 in, 12 out) and 25 N (1 + 4 in, 20 out)
    view = None
    if rank == 0 and args.config == "K":
        L0, R0 = pairs[0]
        engine.process(L0, R0)                      # single-frame path: leaves D1 in the slot's own buffers
        ms_c, ms_r = engine.time_view(iters=50)
        n_px = W * H
        view = {"colormap_us": round(ms_c * 1e3, 2), "colormap_gbs": round(16 * n_px / (ms_c * 1e-3) / 1e9, 1),
                "reproject_us": round(ms_r * 1e3, 2), "reproject_gbs": round(25 * n_px / (ms_r * 1e-3) / 1e9, 1),
                "note": "L2-resident at this size (7-12 MB per launch)"}
    # the feature filters of libviso2's Matcher on a device-resident 1248x375 image:
    # algorithmic bytes 7 N (1 in, du + dv + two int16 maps out)
    filters = None
    if rank == 0 and args.config == "K":
        import ctypes as C
        lib = elas_b200.load_library()
        d_img = d_I[0, 0].contiguous()
        d_du, d_dv = torch.empty_like(d_img), torch.empty_like(d_img)
        d_f1 = torch.empty(d_img.shape, dtype=torch.int16, device=dev); d_f2 = torch.empty_like(d_f1)
        ms_f = C.c_float(0)
        for it in (3, 50):
            rc_f = lib.elas_b200_matcher_filters(local_rank, d_img.data_ptr(), bpl, H, d_du.data_ptr(), d_dv.data_ptr(),
                                                 d_f1.data_ptr(), d_f2.data_ptr(), it, C.byref(ms_f))
        if rc_f == 0:
            filters = {"us": round(ms_f.value * 1e3, 2), "gbs": round(7 * bpl * H / (ms_f.value * 1e-3) / 1e9, 1),
                       "note": "sobel5x5 + blob5x5 + checkerboard5x5 fused, one launch per image, L2-resident at this size"}
    # the synchronous drop-in call itself (what Elas::process forwards to, one frame at a time, pageable host
    # buffers, nothing pipelined): the latency-bound number a caller like StereoThread::run sees
    drop_in = None
    if rank == 0 and args.config == "K":
        import ctypes as C
        lib = elas_b200.load_library()
        o1 = np.empty((H, W), np.float32); o2 = np.empty((H, W), np.float32)
        dims = (C.c_int32 * 3)(W, H, W)
        call = lambda Lx, Rx: lib.elas_b200_process(C.byref(params), Lx.ctypes.data, Rx.ctypes.data, o1.ctypes.data, o2.ctypes.data, dims)
        assert call(*pairs[0]) == 0                             # creates and caches the single-slot context
        t0 = time.perf_counter()
        n_calls = 40
        for i in range(n_calls):
            call(*pairs[i % args.distinct])
        dt = time.perf_counter() - t0
        drop_in = {"value": round(n_calls / dt, 1), "unit": "pairs/s", "ms_per_call": round(1e3 * dt / n_calls, 3),
                   "note": "elas_b200_process, synchronous, one frame in flight, pageable numpy buffers"}
    engine_frames, engine_mesh = engine.frames_per_group, engine.mesh_on_device
    engine.close()

    # the bandwidth-ceiling configuration (BASELINE.json configs[4] geometry, one GPU): its working set
    # (683 MB algorithmic) does not fit L2, so this is the honest HBM-roofline point of the same kernel
    roof_4k = roof_hd = roof_k16 = None
    if rank == 0 and world == 1 and not args.no_4k and args.config == "K":
        def roof_point(Wx, Hx, Dx, frames=0):
            Lx, Rx, _ = synth.synthetic_pair(Wx, Hx, Dx, seed=0)
            ex = elas_b200.ElasB200(elas_b200.stereomapper(Dx), Wx, Hx, n_slots=1, device=local_rank, frames_per_group=frames)
            nfx = ex.frames_per_group
            ex.process_batch([Lx] * nfx, [Rx] * nfx)
            msx, nf = ex.time_matching(iters=20, flush_l2=True, per_frame=False)
            ex.close()
            bx = algorithmic_bytes_matching(Wx, Hx, Dx) * nf
            ax = bx / (msx * 1e-3) / 1e9
            return {"workload": f"synthetic {Wx}x{Hx}, d_max={Dx}", "bound": "hbm", "achieved": round(ax, 1), "peak": peak,
                    "unit": "GB/s", "frac": round(ax / peak, 4), "algorithmic_bytes_per_launch": bx,
                    "frames_per_launch": nf, "ms_per_launch": round(msx, 5)}
        roof_hd = roof_point(1920, 1080, 128)       # BASELINE.json configs[2] geometry
        roof_4k = roof_point(4096, 2160, 256)       # BASELINE.json configs[4] geometry
        # the same kernel on the metric's workload with 16 frames per launch (20 waves of CTAs instead of 10): what
        # the wave quantisation of a launch costs; the pipeline keeps 8 (same pairs/s, finer-grained copies)
        roof_k16 = roof_point(W, H, DMAX, frames=16)
        roof_k16["note"] = "K with 16 frames per launch (opt-in frames_per_group=16); the pipeline and `roofline` use 8"
        roof_4k["traffic"] = ncu_traffic_k7("bandwidth_config")

    (launches,) = sharding.sum_over_ranks([launches], dev)      # whole job

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            arm = CpuArm(args.config)
            per_core = args.cpu_pairs_per_core or CONFIGS[args.config][4]
            rate, dt = arm.run(per_core)
            cpu = {"value": round(rate, 3), "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind,
                   "sample": f"{arm.cores} processes x {per_core} pairs, Elas::process only, {dt:.1f} s wall"}
            arm.close()
        total_pairs = world * B * args.steps
        line = {
            "metric": METRIC, "value": round(total_pairs / (ms_dev_max * 1e-3), 2), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_dev_max / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "frame_groups_per_gpu": slots,
                       "frames_per_launch_chain": engine_frames, "host_workers_per_gpu": workers,
                       "mesh_stage": "device" if engine_mesh else "host",
                       "distinct_pairs": args.distinct, "l2": "512 MiB buffer rewritten between timed steps",
                       "parallelism": f"frame-sharded x{world}, one NCCL broadcast of the parameter block",
                       "host_cores": cores},
            "e2e": {"value": round(total_pairs / (ms_host_max * 1e-3), 2), "unit": "pairs/s",
                    # whole job, like `value`
                    "h2d_bytes_per_step": world * B * 2 * W * H,
                    # D1 as float32; D2 (final after the L/R check: integers or -10) crosses narrowed -- one byte per
                    # pixel plus a validity bit per pixel when disp_max <= 255, int16 otherwise -- and is widened into the
                    # caller's float map by the library (elas_b200.cu)
                    "d2h_bytes_per_step": world * B * d2h_bytes_per_pair,
                    "ms_per_step": round(ms_host_max / args.steps, 4),
                    "bound": f"PCIe device->host: {d2h_bytes_per_pair / 1e6:.2f} MB per pair (D1 float32 + D2 narrowed)"},
            "e2e_left_map_only": {"value": round(world * B / (ms_left_max * 1e-3), 2), "unit": "pairs/s",
                                  "d2h_bytes_per_step": world * B * W * H * 4,
                                  "note": "opt-in D2 == NULL: only the left map returns (what stereomapper reads, stereothread.cpp:116-147)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_matching (K7, left+right, one launch per frame group)", "frames_per_launch": k7_frames,
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": ncu_traffic_k7(frames_per_launch=k7_frames) if args.config == "K" else None,
                         "algorithmic_bytes_per_launch": b_match, "ms_per_launch": round(k7_ms, 5),
                         "peak_source": peak_src},
            "roofline_bandwidth_config": roof_4k,
            "roofline_hd_config": roof_hd, "roofline_16_frames_per_launch": roof_k16,
            "view_kernels": view, "matcher_filters": filters,
            "drop_in_call": drop_in,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "checks": {"frames_not_ok": bad_dev + bad_host + bad_left, "device_and_host_paths_identical": same,
                       "oracle_frames_checked": oracle_checked, "oracle_mismatch": oracle_mismatch},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
