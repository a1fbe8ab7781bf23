"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(brian-team/brian2, `cpp_standalone` device, serial, strict floating-point flags -- SURVEY.md
section 8c) on the model scripts of tests/models.py.

Run in the build container (needs baseline/_ref, i.e. /root/reference):
    python tests/golden/make_golden.py [case ...]
The resulting .npz files are committed; the GPU box never needs the reference to check parity.
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

#: case name -> (model, kwargs).  Sizes are chosen so that every fixture stays small.
CASES = {
    "cuba_4000": ("cuba", dict(N=4000, p=0.02, duration=0.2)),
    "cuba_1000": ("cuba", dict(N=1000, p=0.08, duration=0.1)),
    "cobahh_1000": ("cobahh", dict(N=1000, duration=0.05)),
    # one biological second of the Hodgkin-Huxley network: how long does the device (CUDA exp /
    # expm1, <= 1-2 ulp from glibc) stay spike-exact?  (spike train only)
    "cobahh_1000_long": ("cobahh", dict(N=1000, duration=1.0, trace=())),
    # one biological second of COBAHH-4000 for `prefs.devices.b200.libm = 'glibc'` (device exp /
    # expm1 / pow with glibc's arithmetic): final state of every neuron, spike counts and a digest
    # of the spike train (136 k spikes) -- everything must be bit-identical
    "cobahh_4000_1s": ("cobahh", dict(N=4000, duration=1.0, trace=())),
    "brunel_hetero": ("brunel", dict(N_E=800, epsilon=0.1, duration=0.1, hetero_delays=True)),
    "brunel_homog": ("brunel", dict(N_E=800, epsilon=0.1, duration=0.1, hetero_delays=False)),
    "stdp_1000": ("stdp", dict(N=1000, duration=0.2)),
    "synapses_only": ("synapses_only", dict(N=2000, p=0.2, rate_hz=100.0, duration=0.005)),
    "synapses_only_delay": ("synapses_only", dict(N=2000, p=0.2, rate_hz=100.0, duration=0.005, delay_steps=3)),
    # heavy steps: > 64 lines of 32 synapses per warp of the grid -> the ticket-based (dynamic)
    # distribution of the propagation kernel
    "synapses_only_heavy": ("synapses_only", dict(N=60000, p=0.3, rate_hz=100.0, duration=0.0005)),
    # many short rows per step (6000 spiking sources x ~20 synapses, 5 delay bins): the gather mode
    # of the propagation kernel
    "synapses_only_short": ("synapses_only", dict(N=1000, p=0.02, rate_hz=60000.0, duration=0.002,
                                                   hetero_bins=5)),
    # Potjans-Diesmann microcircuit at 1 % of the neurons (in-degrees preserved), DC background
    "potjans_small": ("potjans", dict(scale=0.01, duration=0.05)),
    "submon": ("submon", dict(N=600, duration=0.05)),
    "ragged": ("ragged", dict(N=600, duration=0.03)),
    "spikegen": ("spikegen", dict(N=200, n_spikes=3000, duration=0.05)),
    "spikegen_period": ("spikegen", dict(N=200, n_spikes=600, duration=0.05, period_ms=10.0)),
    "gapjunction": ("gapjunction", dict(N=300, p=0.1, duration=0.05)),
    "timedarray": ("timedarray", dict(N=100, duration=0.05)),
    # stochastic (in-loop RNG): the golden file holds the reference's statistics only
    "poisson_drive": ("poisson_drive", dict(N=2000, duration=0.1)),
    "poissonfn": ("poissonfn", dict(N=4000, duration=0.02)),
    # several clocks + scalar writes, shared variable on the main clock,
    # clock-driven synaptic equations + summed variable
    "multiclock": ("multiclock", dict(N=200, duration=0.05)),
    "sharedvar": ("sharedvar", dict(N=300, duration=0.03)),
    "synstate": ("synstate", dict(N=150, duration=0.03)),
    # exp / expm1 / exprel / log / pow (+ the powers g++ folds) per neuron and step over wide,
    # drifting argument ranges: bit-identical with prefs.devices.b200.libm = 'glibc'
    "mathfuncs": ("mathfuncs", dict(N=2048, duration=0.01)),
}


#: cases that draw random numbers inside the time loop (compared statistically)
STOCHASTIC = {"poisson_drive", "poissonfn"}


def reduce_train(res):
    """Replace the spike train (i, t) by its length and SHA-256 digests (small fixture)."""
    import hashlib

    out = {k: v for k, v in res.items() if k not in ("spikes_i", "spikes_t")}
    out["spikes_n"] = np.array([len(res["spikes_i"])], dtype=np.int64)
    for key in ("spikes_i", "spikes_t"):
        digest = hashlib.sha256(np.ascontiguousarray(res[key]).tobytes()).digest()
        out[key + "_sha256"] = np.frombuffer(digest, dtype=np.uint8).copy()
    return out


def main(argv):
    names = argv or list(CASES)
    for case in names:
        model, kwds = CASES[case]
        d = tempfile.mkdtemp(prefix=f"golden_{case}_")
        objs, res = models.run_model(b, model, "cpp_standalone", d, **kwds)
        res = {k: v for k, v in res.items() if k != "last_run_time"}
        if case.endswith("_long"):
            res = {k: v for k, v in res.items() if k in ("spikes_i", "spikes_t")}
        if case.endswith("_1s"):   # state + counts + digest of the train instead of the train
            res = reduce_train(res)
        if case in STOCHASTIC:   # statistics only
            res = {k: v for k, v in res.items() if not k.endswith(("_i", "_t"))}
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **res)
        summary = {k: (v.shape, str(v.dtype)) for k, v in res.items()}
        print(case, os.path.getsize(path), "bytes", summary)


if __name__ == "__main__":
    main(sys.argv[1:])
