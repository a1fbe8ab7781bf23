"""Debug helper: run one golden case on the b200 device and report per-key differences."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import brian2_b200  # noqa
import brian2 as b

import models
from golden.make_golden import CASES

for case in sys.argv[1:]:
    model, kwds = CASES[case]
    d = os.path.join(ROOT, "brian2_b200", "_prebuilt", "debug_" + case)
    prefs_update = {}
    if os.environ.get("B200_STEPWISE"):
        prefs_update["devices.b200.persistent"] = False
    objs, res = models.run_model(b, model, "b200", d, prefs_update=prefs_update, **kwds)
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))
    print("==", case, "loop time", b.device._last_run_time, "events", b.device.counter("events"),
          "launches", b.device.counter("launches"))
    for key in gold.files:
        g, r = gold[key], res[key]
        if g.shape != r.shape:
            print(f"  {key}: SHAPE {r.shape} vs golden {g.shape}")
            n = min(len(g), len(r))
            if n and g.ndim == 1:
                neq = np.nonzero(g[:n] != r[:n])[0]
                print("     first mismatch at", neq[:5], "of", n, g[neq[:5]], r[neq[:5]])
            continue
        if np.array_equal(g, r):
            print(f"  {key}: identical")
        else:
            neq = np.nonzero(g != r)
            rel = np.max(np.abs(g - r) / (np.abs(g) + 1e-300)) if g.dtype.kind == "f" else -1
            print(f"  {key}: {len(neq[0])} of {g.size} differ, max rel {rel:.3e}, first at {neq[0][:5]}: gold {g[neq][:5]} got {r[neq][:5]}")
