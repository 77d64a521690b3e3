"""Map of log10(w_right) -- the proposal mass of the alternate sampler's right (truncated gamma)
piece -- over (h, |z|/2), computed with the CPU oracle's setup code.  The CUDA path treats
elements with |z| >= 14 as left-only (`alternate_left_only` in pgm_methods.cuh): this script
prints the map and the largest weight inside that region (must stay below 2^-33)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import oracle as O  # noqa: E402

hs = [0.01, 0.1, 0.3, 0.5, 0.8, 1, 1.06, 1.125, 1.25, 1.5, 2, 2.5, 3, 3.5, 4]
zs = [0, 0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 5, 6, 7, 8, 10, 15, 25]
print("log10(w_right); rows h, columns |z|/2")
print("      " + " ".join(f"{z:6.1f}" for z in zs))
worst = -999.0
for h in hs:
    row = []
    for zz in zs:
        wr = O.setup_probe("alternate", h, 2 * zz)[0]
        lw = np.log10(wr) if wr > 0 else -999.0
        row.append(lw)
        if zz >= 7:
            worst = max(worst, lw)
    print(f"{h:5.2f} " + " ".join(f"{max(v, -99):6.1f}" for v in row))
print(f"largest log10(w_right) for |z| >= 14: {worst:.2f} (2^-33 = {np.log10(2.0**-33):.2f})")
