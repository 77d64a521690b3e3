"""Map of log10(w_right) -- the proposal mass of the saddle sampler's right (truncated gamma)
envelope piece -- over (h, |z|/2), computed with the CPU oracle's setup code.

The CUDA path specialises saddle elements with w_right < 2^-33 (the 32-bit mixture uniform can
never select the piece): `saddle_left_only(h, z)` in polyagamma_b200/csrc/pgm_methods.cuh is
the rule  h * (1 + |z| / 8) >= 22.  This script prints the map and checks the rule against it.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import oracle as O  # noqa: E402

hs = [4.01, 4.5, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 22, 25, 30, 50]
zs = [0, 0.25, 0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 5, 6, 8, 10, 15, 25]
print("log10(w_right); rows h, columns |z|/2")
print("      " + " ".join(f"{z:6.2f}" for z in zs))
worst = -99.0
for h in hs:
    row = []
    for zz in zs:
        wr = 1.0 - O.setup_probe("saddle", h, 2 * zz)[0]
        lw = np.log10(wr) if wr > 0 else -99.0
        row.append(lw)
        if h * (1 + 0.25 * zz) >= 22.0:
            worst = max(worst, lw)
    print(f"{h:5.2f} " + " ".join(f"{v:6.1f}" for v in row))
print(f"largest log10(w_right) inside the left-only region: {worst:.2f} (2^-33 = {np.log10(2.0**-33):.2f})")
