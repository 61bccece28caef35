"""Writes tests/golden/view_small.npz from the reference's own statements (oracle/_ref/libview_ref.so,
built by oracle/Makefile from stereomapper/stereothread.cpp:116-147 and :180-255).  Run in the build
container, where /root/reference exists:  python tests/golden/view/make_view_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers  # noqa: E402
from view_cases import view_case  # noqa: E402

I1, D1, view, H = view_case("odd_small")
ref = checkers.ViewChecker("ref")
I, D, X, Y, Z = ref.reproject(I1, D1, view, H)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "view", "view_small.npz"), I1=np.ascontiguousarray(I1), D1=D1, view=view, H=H,
                    color=ref.colormap(D1), I=I, D=D, X=X, Y=Y, Z=Z)
print("wrote view_small.npz")
