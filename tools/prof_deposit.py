"""Batched 1-D deposit on the pair-RDF grid, alone, for ncu captures of deposit1d_owner_kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))
import edm_b200 as edm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = edm.GaussGrid(1, [1.68], [5.0], [0.00025], [0], 1, [0.025])
rng = np.random.default_rng(99)
c = rng.uniform(1.68, 5.0, n)
h = np.full(n, 1e-6)
for _ in range(2):
    ba = g.add_values(c, h)
print("deposited", n, "hills twice; integral of the last batch", float(np.sum(ba)))
