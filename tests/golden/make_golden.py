"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libelas_ref.so).

Run here (the container that has /root/reference):  python tests/golden/make_golden.py
The reference holds no golden vectors of its own for Elas::process, so these are
outputs of the reference itself, committed so that the GPU box (which has no /root/reference) can
pin both the oracle restatement and the CUDA path.

Cases:
  synth_320x120_d63   seeded synthetic pair (regenerated from the seed, inputs not stored)
  urban1_crop         a 480x160 crop of libelas/img/urban1 (inputs stored, it is reference DATA)
  full/urban{1..4}    the four 1344x391 street-scene pairs of libelas/img at full size (the reference's
                      KITTI-size real images, main.cpp:105-113), stereomapper's parameter set, d_max 255:
                      inputs, integer stages, raw / final maps (integer-valued maps stored as int16)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "stereo-vision_b200"))
import checkers  # noqa: E402
import synth  # noqa: E402

KEEP = ["dcan_raw", "dcan", "lattice_dims", "support", "tri1", "tri2", "planes1", "planes2",
        "D1_raw", "D2_raw", "D1_lr", "D2_lr", "D1_seg", "D1_gap", "D1", "D2"]


def cases():
    L, R, _ = synth.synthetic_pair(320, 120, 63, seed=3)
    yield "synth_320x120_d63", L, R, checkers.stereomapper(63), False
    yield "synth_320x120_d63_demo", L, R, checkers.demo(63), False
    img = "/root/reference/libelas/img/urban1_%s.pgm"
    l = synth.read_pgm(img % "left")[150:310, 400:880]
    r = synth.read_pgm(img % "right")[150:310, 400:880]
    yield "urban1_crop", np.ascontiguousarray(l), np.ascontiguousarray(r), checkers.stereomapper(127), True


FULL_INT16 = ["D1_raw", "D2_raw", "D2"]          # integer-valued maps (or -10 / -1): exact as int16
FULL_KEEP = ["dcan", "support", "tri1", "tri2", "D1"]


def full_cases():
    for k in (1, 2, 3, 4):
        img = "/root/reference/libelas/img/urban%d_%%s.pgm" % k
        yield "urban%d" % k, synth.read_pgm(img % "left"), synth.read_pgm(img % "right"), checkers.stereomapper(255)


def main_full(ref):
    os.makedirs(os.path.join(HERE, "full"), exist_ok=True)
    for name, L, R, p in full_cases():
        rc, D1, D2, st = ref.run_stages(L, R, p)
        rc2, E1, E2 = ref.process(L, R, p)
        assert rc == 0 and np.array_equal(D1, E1) and np.array_equal(D2, E2)
        out = {k: st[k] for k in FULL_KEEP}
        for k in FULL_INT16:
            assert np.array_equal(st[k], st[k].astype(np.int16).astype(np.float32))
            out[k] = st[k].astype(np.int16)
        out["params"] = np.frombuffer(bytes(p), np.uint8)
        out["I1"], out["I2"] = L, R
        path = os.path.join(HERE, "full", name + ".npz")
        np.savez_compressed(path, **out)
        print("full/" + name, "support", len(st["support"]) // 3, "tri", len(st["tri1"]) // 3, len(st["tri2"]) // 3,
              "valid D1 %.3f" % (D1 >= 0).mean(), os.path.getsize(path), "bytes")


def main():
    checkers.build("ref")
    ref = checkers.RefElas()
    main_full(ref)
    for name, L, R, p, store_inputs in cases():
        rc, D1, D2, st = ref.run_stages(L, R, p)
        rc2, E1, E2 = ref.process(L, R, p)
        assert rc == 0 and np.array_equal(D1, E1) and np.array_equal(D2, E2)
        out = {k: st[k] for k in KEEP}
        out["params"] = np.frombuffer(bytes(p), np.uint8)
        if store_inputs:
            out["I1"], out["I2"] = L, R
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "support", len(st["support"]) // 3, "tri", len(st["tri1"]) // 3,
              "valid D1 %.3f" % (D1 >= 0).mean(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
