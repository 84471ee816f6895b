"""TEST INFRASTRUCTURE (container-only): golden history of the reference's random-field test.

`integration_tests/integration_test.py:376-421` (Test1DWavePropagation_2D.test_2): quad4 column, heaviside load, Young's
modulus from a gstools Exponential random field (seed 26021981), compared by the reference against
`integration_tests/results_rf_2d/data.pickle`.  That pickle is copied here as arrays -> tests/golden/history_rf_2d.npz.
It pins the restatement of gstools' RandMeth mode sampler in scatter_b200/random_fields.py (`gstools_modes`): the
history only comes out right if every element gets the Young's modulus gstools 1.7.0 gave it.

    python oracle/make_golden_rf2d.py
"""
import os
import pickle

import numpy as np

IT = "/root/reference/integration_tests"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

with open(os.path.join(IT, "results_rf_2d/data.pickle"), "rb") as f:
    g = pickle.load(f)
ids = list(g["nodes"])
h = {"nodes": np.array(ids), "time": np.asarray(g["time"]), "position": np.asarray(g["position"])}
for name in ("displacement", "velocity", "acceleration"):
    h[name] = np.array([[g[name][str(n)][lab] for lab in "xy"] for n in ids])       # (nn, 2, nt)
np.savez_compressed(os.path.join(OUT, "history_rf_2d.npz"), **h)
print("history_rf_2d.npz", os.path.getsize(os.path.join(OUT, "history_rf_2d.npz")) // 1024, "KiB")
