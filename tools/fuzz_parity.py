"""Randomised differential test of the CUDA path against the CPU oracle: random image sizes, disparity ranges and
parameter blocks (ROBOTICS family, MIDDLEBURY preset, subsampling, odd grid sizes, lattice steps, thresholds), single
frames and batches over frame groups.  Every stage of the single-frame run and the final maps of the batch run are
compared bit for bit.  Usage: python tools/fuzz_parity.py [cases] [seed] [max_w] [max_h]   (ELAS_B200_HOST_STAGE=1 forces the host mesh stage)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import checkers, elas_b200, synth

STAGES = ["desc1", "desc2", "dcan_raw", "dcan", "support", "tri1", "tri2", "grid1", "grid2", "D1_raw", "D2_raw", "D1_lr", "D2_lr",
          "D1_seg", "D1_gap", "D1_mean", "D1", "D2"]


def run_cases(cases, seed=1, kinds=("stereomapper", "demo", "middlebury", "sub", "odd"), max_w=720, max_h=280):
    """Runs `cases` random cases; returns the list of (description, differing stages)."""
    rng = np.random.default_rng(seed)
    oracle = checkers.OracleElas()
    failures = []
    bad = 0
    t0 = time.time()
    for case in range(cases):
        dmax = int(rng.choice([31, 47, 63, 95, 100, 127, 200, 255]))
        W = int(rng.integers(max(96, dmax + 40), max(max_w, dmax + 60))); H = int(rng.integers(64, max_h))
        kind = rng.choice(list(kinds))
        p = checkers.stereomapper(dmax) if kind in ("stereomapper", "sub", "odd") else checkers.demo(dmax) if kind == "demo" else checkers.middlebury().copy(disp_max=dmax)
        over = {}
        if kind == "sub": over["subsampling"] = 1
        if kind == "odd":
            over.update(grid_size=int(rng.choice([10, 15, 16, 20, 24, 25])), candidate_stepsize=int(rng.choice([4, 5, 6, 7])),
                        speckle_size=int(rng.choice([20, 100, 200, 400])), lr_threshold=int(rng.choice([1, 2])),
                        ipol_gap_width=int(rng.choice([0, 1, 3, 7])), match_texture=int(rng.choice([0, 1, 30])),
                        support_texture=int(rng.choice([5, 10, 30])), incon_min_support=int(rng.choice([3, 5, 8])),
                        filter_adaptive_mean=int(rng.integers(0, 2)), postprocess_only_left=int(rng.integers(0, 2)))
            if rng.random() < 0.3: over["disp_min"] = int(rng.integers(1, 6))
            if rng.random() < 0.3: over["sradius"] = 3.0
            if rng.random() < 0.25: over["sigma"] = float(rng.choice([0.6, 1.5, 2.0]))        # plane radius 2..6: the generic window
            if rng.random() < 0.25: over.update(gamma=float(rng.choice([1.0, 5.0, 15.0])), beta=float(rng.choice([0.01, 0.05])))
            if rng.random() < 0.2: over["filter_median"] = 1
            if rng.random() < 0.2: over["support_threshold"] = float(rng.choice([0.7, 0.95]))
            if rng.random() < 0.15: over["subsampling"] = 1
            if rng.random() < 0.15: over["add_corners"] = 1
            if rng.random() < 0.2: over["incon_window_size"] = int(rng.choice([3, 7]))
            if rng.random() < 0.2: over["speckle_sim_threshold"] = float(rng.choice([0.5, 2.0]))
        p = p.copy(**over)
        H = max(H, 2 * p.grid_size + 8)
        L, R, _ = synth.synthetic_pair(W, H, dmax, seed=int(rng.integers(0, 1 << 30)))
        if rng.random() < 0.2:                                        # a textureless band / noise to provoke holes and speckles
            L = L.copy(); L[H // 3: H // 3 + 12] = 128
        content = rng.random()
        if content < 0.08:                                            # pure noise: hardly any support point survives
            L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = rng.integers(0, 256, (H, W), dtype=np.uint8)
        elif content < 0.14:                                          # the whole scene at (nearly) the largest disparity
            tex = rng.integers(0, 256, (H, W + dmax), dtype=np.uint8)
            R = tex[:, dmax:].copy(); L = tex[:, 1:W + 1].copy() if rng.random() < 0.5 else tex[:, :W].copy()
        elif content < 0.2:                                           # salt noise on the right image: speckles and failed L/R checks
            R = R.copy(); m = rng.random((H, W)) < 0.05; R[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
        if rng.random() < 0.3:                                        # the caller's rows are wider than the image (stereothread.cpp:111)
            pitch = W + int(rng.integers(1, 40))
            Lp = rng.integers(0, 256, (H, pitch), dtype=np.uint8); Rp = rng.integers(0, 256, (H, pitch), dtype=np.uint8)
            Lp[:, :W] = L; Rp[:, :W] = R
            L, R = Lp[:, :W], Rp[:, :W]                                # views with strides[0] = pitch
        tag = f"case {case}: {W}x{H} d{dmax} {kind} {over}"
        if os.environ.get("FUZZ_VERBOSE"): print("RUN", tag, "pitch", L.strides[0], flush=True)
        try:
            rc_o, O1, O2, st_o = oracle.run_stages(np.ascontiguousarray(L), np.ascontiguousarray(R), p)
            pp = elas_b200.Params.from_buffer_copy(bytes(p))
            e = elas_b200.ElasB200(pp, W, H, n_slots=2, n_workers=2, frames_per_group=int(rng.choice([1, 2, 3, 8])))
            try:
                rc, D1, D2 = e.process(L, R, capture=True)
                diffs = []
                if rc != rc_o: diffs.append(f"rc {rc} vs {rc_o}")
                if rc_o == 0:
                    for k in STAGES:
                        a = e.stage(k); b = st_o.get(k)
                        if a is None or b is None: continue
                        if a.shape != b.shape or not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
                            diffs.append(k)
                            if a.shape == b.shape and k.endswith("_raw"):
                                aa, bb = a.reshape(D1.shape), b.reshape(D1.shape)
                                w = np.argwhere(aa != bb)
                                print(f"   {k}: {len(w)} pixels differ; first:", [(int(v), int(u), float(aa[v, u]), float(bb[v, u])) for v, u in w[:6]], flush=True)
                    if not (np.array_equal(D1.view(np.uint32), O1.view(np.uint32)) and np.array_equal(D2.view(np.uint32), O2.view(np.uint32))): diffs.append("final")
                    n = int(rng.integers(1, 7))
                    for rep in range(2):
                        st, B1, B2 = e.process_batch([L] * n, [R] * n)
                        for i in range(n):
                            if st[i] != 0 or not (np.array_equal(B1[i].view(np.uint32), O1.view(np.uint32)) and np.array_equal(B2[i].view(np.uint32), O2.view(np.uint32))):
                                diffs.append(f"batch rep {rep} frame {i}/{n}")
                        n = int(rng.integers(1, 7))
                    if rng.random() < 0.35:
                        # a batch of DIFFERENT frames (one of them blank), in random order, through host or device buffers
                        Lc, Rc = np.ascontiguousarray(L), np.ascontiguousarray(R)
                        L2, R2, _ = synth.synthetic_pair(W, H, dmax, seed=int(rng.integers(0, 1 << 30)))
                        blank = np.full((H, W), 77, np.uint8)
                        frames = [(Lc, Rc, rc_o, O1, O2)]
                        rc2, P1, P2 = oracle.process(L2, R2, p)
                        frames.append((L2, R2, rc2, P1, P2))
                        # a blank frame: fewer than 3 support points (rc 1, maps -10) -- unless add_corners invents six (elas.cpp:283-318)
                        rc3, Q1, Q2 = oracle.process(blank, blank, p)
                        frames.append((blank, blank, rc3, Q1, Q2) if rc3 == 0 else (blank, blank, 1, np.full_like(O1, -10), np.full_like(O2, -10)))
                        order = [int(x) for x in rng.integers(0, 3, int(rng.integers(2, 12)))]
                        if rng.random() < 0.5:
                            import torch
                            dI = torch.stack([torch.stack([torch.from_numpy(frames[k][0]), torch.from_numpy(frames[k][1])]) for k in order]).cuda()
                            dD = torch.full((len(order), 2) + O1.shape, -77.0, dtype=torch.float32, device="cuda")
                            st = e.process_batch_ptrs([dI[i, 0].data_ptr() for i in range(len(order))], [dI[i, 1].data_ptr() for i in range(len(order))],
                                                      [dD[i, 0].data_ptr() for i in range(len(order))], [dD[i, 1].data_ptr() for i in range(len(order))], W, device=True)
                            torch.cuda.synchronize()
                            outs = dD.cpu().numpy(); B1 = [outs[i, 0] for i in range(len(order))]; B2 = [outs[i, 1] for i in range(len(order))]
                            how = "device"
                        elif rng.random() < 0.5:
                            # pinned host buffers, some frames without a right map (D2 == NULL)
                            import torch
                            hI = torch.stack([torch.stack([torch.from_numpy(frames[k][0]), torch.from_numpy(frames[k][1])]) for k in order]).pin_memory()
                            hD = torch.full((len(order), 2) + O1.shape, -77.0, dtype=torch.float32).pin_memory()
                            skip2 = [bool(rng.random() < 0.4) for _ in order]
                            st = e.process_batch_ptrs([hI[i, 0].data_ptr() for i in range(len(order))], [hI[i, 1].data_ptr() for i in range(len(order))],
                                                      [hD[i, 0].data_ptr() for i in range(len(order))],
                                                      [0 if skip2[i] else hD[i, 1].data_ptr() for i in range(len(order))], W, device=False)
                            outs = hD.numpy(); B1 = [outs[i, 0] for i in range(len(order))]
                            B2 = [frames[k][4] if skip2[i] else outs[i, 1] for i, k in enumerate(order)]
                            for i in range(len(order)):
                                if skip2[i] and not (outs[i, 1] == -77.0).all(): diffs.append(f"mixed pinned batch frame {i}: a right map that was not asked for was written")
                            how = "pinned"
                        else:
                            st, B1, B2 = e.process_batch([frames[k][0] for k in order], [frames[k][1] for k in order])
                            how = "host"
                        for i, k in enumerate(order):
                            want_rc = frames[k][2]
                            if st[i] != want_rc or not (np.array_equal(B1[i].view(np.uint32), frames[k][3].view(np.uint32)) and np.array_equal(B2[i].view(np.uint32), frames[k][4].view(np.uint32))):
                                diffs.append(f"mixed {how} batch frame {i} (kind {k}) of {order}: status {st[i]} vs {want_rc}")
            finally:
                e.close()
        except Exception as ex:                                       # unsupported combinations must be refused cleanly, not crash
            diffs = [f"exception {type(ex).__name__}: {ex}"]
        if diffs:
            bad += 1
            failures.append((tag, diffs))
            print("DIFF", tag, "->", diffs[:6], flush=True)
    print(f"fuzz: {cases} cases, {bad} with differences, {time.time() - t0:.0f} s", flush=True)
    return failures


if __name__ == "__main__":
    # python tools/fuzz_parity.py [cases] [seed] [max_w] [max_h]
    a = [int(x) for x in sys.argv[1:]]
    run_cases(a[0] if len(a) > 0 else 40, a[1] if len(a) > 1 else 1, max_w=a[2] if len(a) > 2 else 720, max_h=a[3] if len(a) > 3 else 280)
