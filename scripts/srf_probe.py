"""Time (and, under ncu, profile) the random-field kernel k_srf: Gaussian model, 1000 modes, N points of a box lattice."""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scatter_b200 import _lib, random_fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(0)
pos = rng.uniform(0, 127.5, (n, 3))
sf = random_fields.SpectralField("Gaussian", 3, var=0.0011, mean=17.2, len_scale=[20.0, 2.0, 20.0], angles=0.0, seed=26021981)
ctx = _lib.Context(0)
for _ in range(3):
    out = sf(pos, lognormal=True, ctx=ctx)
    print(f"k_srf: {n} points x {sf.mode_no} modes in {sf.seconds_device * 1e3:.2f} ms -> {n * sf.mode_no / sf.seconds_device / 1e9:.1f} G sincos/s; "
          f"mean {out.mean():.4e} std {out.std():.4e}")
