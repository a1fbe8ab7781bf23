"""Diagnostic (GPU box): COBAHH-1000 for 1 and 10 timesteps on cpp_standalone (strict flags) and
on the b200 device with `prefs.devices.b200.libm = 'glibc'`; prints, per state variable, how many
values are not bit-identical.  After one step every variable isolates its own functions
(h: exp, pow(exp, c); m: exprel; n: exprel, pow(exp, c); v: pow(n, 4), pow(m, 3), exp)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import brian2_b200  # noqa: E402,F401
import brian2 as b  # noqa: E402
import models  # noqa: E402

for steps in (1, 10):
    kw = dict(N=1000, duration=steps * 1e-4, trace=())
    _, ref = models.run_model(b, "cobahh", "cpp_standalone", tempfile.mkdtemp(prefix="diag_ref_"), **kw)
    for libm in ("glibc", "cuda"):
        d = os.path.join(ROOT, "brian2_b200", "_prebuilt", f"diag_{libm}_{steps}")
        try:
            _, dev = models.run_model(b, "cobahh", "b200", d, prefs_update={"devices.b200.libm": libm}, **kw)
        finally:
            b.prefs["devices.b200.libm"] = "cuda"
        for key in ("P_v", "P_m", "P_n", "P_h", "P_ge", "P_gi"):
            bad = int((ref[key].view(np.uint64) != dev[key].view(np.uint64)).sum())
            rel = float(np.max(np.abs(ref[key] - dev[key]) / np.abs(ref[key])))
            print(f"steps={steps} libm={libm} {key}: {bad} of {ref[key].size} differ, max rel {rel:.3g}")
