"""Writes tests/golden/viso/filters_small.npz: the feature maps libviso2's own filter.cpp (compiled by oracle/Makefile
into oracle/_ref/libvisofilter_ref.so) produces for two small seeded images.  Run in the container that holds
/root/reference; the .npz is committed, the GPU box only reads it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers  # noqa: E402
from filter_cases import filter_case  # noqa: E402

checkers.build("all")
ref = checkers.MatcherFilterChecker("ref")
out = {}
for name in ("saturating", "narrow"):
    I = filter_case(name)
    du, dv, f1, f2 = ref(I)
    out.update({f"{name}_I": I, f"{name}_du": du, f"{name}_dv": dv, f"{name}_f1": f1, f"{name}_f2": f2})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "viso", "filters_small.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
