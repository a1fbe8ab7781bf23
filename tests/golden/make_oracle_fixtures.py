"""Fixtures that PIN the CPU restatement (oracle/hotpath_oracle.c): inputs (initial state,
connectivity, delays, weights -- all drawn by the reference's own host code) and outputs (spike
trains, final state) of the UNMODIFIED reference on `cpp_standalone` (serial, strict flags).

Each model is built twice with the same seed: once with duration 0 (captures the inputs the run
starts from) and once with the real duration (outputs).  Run in the build container:
    python tests/golden/make_oracle_fixtures.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from brian2_b200._brian2_path import ensure_brian2_importable  # noqa: E402

ensure_brian2_importable()
import brian2 as b  # noqa: E402

import models  # noqa: E402

CASES = {
    "oracle_cuba_400": ("cuba", dict(N=400, p=0.1, duration=0.1)),
    "oracle_brunel_500": ("brunel", dict(N_E=400, epsilon=0.1, duration=0.1, hetero_delays=True)),
    "oracle_stdp_200": ("stdp", dict(N=200, duration=0.3)),
    "oracle_cobahh_300": ("cobahh", dict(N=300, duration=0.1, n_syn_per_neuron=30.0, trace=())),
}


def _inputs(objs):
    out = {}
    for key, obj in objs.items():
        if isinstance(obj, b.Synapses):
            out[f"in_{key}_pre"] = np.asarray(obj.i[:]).astype(np.int32) + int(getattr(obj.source, "start", 0))
            out[f"in_{key}_post"] = np.asarray(obj.j[:]).astype(np.int32) + int(getattr(obj.target, "start", 0))
            out[f"in_{key}_delay"] = np.asarray(obj.delay_[:]).astype(np.float64)
            if "w" in obj.variables:
                out[f"in_{key}_w"] = np.asarray(obj.w_[:]).astype(np.float64)
    for group, var in objs["state"]:
        if not isinstance(objs[group], b.Synapses):
            out[f"in_{group}_{var}"] = np.asarray(getattr(objs[group], var + "_")[:]).copy()
    if "inp" in objs:
        out["in_inp_rate"] = np.asarray(objs["inp"].rate_[:]).copy()
        out["in_inp_x"] = np.asarray(objs["inp"].x_[:]).copy()
    return out


def main(names=None):
    for case, (model, kwds) in CASES.items():
        if names and case not in names:
            continue
        kw0 = dict(kwds, duration=0.0)
        objs0, _ = models.run_model(b, model, "cpp_standalone", tempfile.mkdtemp(prefix=case), **kw0)
        data = _inputs(objs0)
        objs, res = models.run_model(b, model, "cpp_standalone", tempfile.mkdtemp(prefix=case), **kwds)
        chk = _inputs(objs)
        for k in data:   # same seed -> same connectivity
            if k.endswith(("_pre", "_post", "_delay")):
                assert np.array_equal(data[k], chk[k]), k
        data.update({k: v for k, v in res.items() if k != "last_run_time"})
        data["duration"] = np.array([kwds["duration"]])
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **data)
        print(case, os.path.getsize(path), "bytes", sorted(data))


if __name__ == "__main__":
    main(sys.argv[1:])
