"""Synapse creation on the device ("sharded construction").

The reference draws a connect() call from one sequential mt19937 stream
(templates/synapses_create_generator.cpp:155-173); the device version gives every source row its
own Philox stream (csrc/b200_connect.cuh), so the two can only agree STATISTICALLY.  What is
exact: the device result does not depend on the number of GPUs (test_sharded_world_independence,
2 GPUs).  The bit-exact oracle for connectivity remains the default host path
(prefs.devices.b200.construction = 'reference', tests/test_parity_gpu.py)."""
import os

import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run_sharded(brian, project_dir, model, **kwds):
    try:
        return models.run_model(brian, model, "b200", project_dir,
                                prefs_update={"devices.b200.construction": "sharded"}, **kwds)
    finally:
        brian.prefs["devices.b200.construction"] = "reference"


def _binomial_ok(count, n, p, sigmas=5.0):
    return abs(count - n * p) <= sigmas * np.sqrt(n * p * (1 - p)) + 1


def test_connect_p_statistics(brian, project_dir):
    """connect(p=...) with the jump algorithm (p < 0.25) + per-synapse random delays: synapse count,
    in- and out-degree distributions, ordering, delay histogram; and the network it produces
    fires like the reference's (population rate of the golden Brunel run)."""
    N_E, eps = 4000, 0.05
    objs, res = _run_sharded(brian, project_dir, "brunel", N_E=N_E, epsilon=eps, duration=0.05)
    N = N_E + N_E // 4
    for name, n_pre in (("exc", N_E), ("inh", N_E // 4)):
        S = objs[name]
        i, j = np.asarray(S.i[:]), np.asarray(S.j[:])
        assert len(i) == res[f"{name}_nsyn"][0]
        assert _binomial_ok(len(i), n_pre * N, eps), (name, len(i))
        assert i.min() >= 0 and i.max() < n_pre and j.min() >= 0 and j.max() < N
        key = i.astype(np.int64) * N + j
        assert np.all(np.diff(key) > 0), "synapses must be sorted by (pre, post) without duplicates"
        out_deg = np.bincount(i, minlength=n_pre)
        in_deg = np.bincount(j, minlength=N)
        # (the counters are indexed by the ABSOLUTE index in the parent group, like the reference's:
        # synapses_create_generator.cpp:22-23 sizes them N + offset)
        assert np.array_equal(out_deg, np.asarray(S.N_outgoing_pre[:])[-n_pre:])
        assert np.array_equal(in_deg, np.asarray(S.N_incoming_post[:])[-N:])
        # binomial degrees: mean and variance (variance of a sample variance ~ 2 sigma^4 / n)
        for deg, n_other, n_rows in ((out_deg, N, n_pre), (in_deg, n_pre, N)):
            mean, var = n_other * eps, n_other * eps * (1 - eps)
            assert abs(deg.mean() - mean) < 5 * np.sqrt(var / n_rows)
            assert abs(deg.var() - var) < 6 * var * np.sqrt(2.0 / n_rows)
        # gaps between consecutive targets of a row are geometric: P(gap = 1) = p
        same_row = i[1:] == i[:-1]
        gaps = (j[1:] - j[:-1])[same_row]
        assert _binomial_ok(int(np.sum(gaps == 1)), len(gaps), eps)
        # delays '(1 + int(rand()*20)) * 0.1*ms': uniform over 20 values, independent of position
        steps = np.rint(np.asarray(S.delay_[:]) / 1e-4).astype(int)
        assert steps.min() == 1 and steps.max() == 20
        hist = np.bincount(steps, minlength=21)[1:]
        chi2 = np.sum((hist - len(steps) / 20.0) ** 2 / (len(steps) / 20.0))
        assert chi2 < 19 + 6 * np.sqrt(2 * 19), chi2
        assert abs(np.corrcoef(steps[:-1], steps[1:])[0, 1]) < 5.0 / np.sqrt(len(steps))
    # dynamics: same parameters as the golden run (different size and connectivity): the
    # population rate over the run is that of the reference within 25 %
    gold = np.load(os.path.join(GOLDEN, "brunel_hetero.npz"))
    rate_ref = len(gold["spikes_i"]) / 1000.0 / 0.1
    rate = len(res["spikes_i"]) / float(N) / 0.05
    assert 0.75 * rate_ref < rate < 1.25 * rate_ref, (rate, rate_ref)
    assert brian.device.counter("connect_synapses") == res["exc_nsyn"][0] + res["inh_nsyn"][0]


def test_connect_dense_and_conditions(brian, project_dir):
    """p >= 0.25 (one Bernoulli test per candidate), a condition on (i, j), subgroup offsets and
    `range` generators: exact where the result is deterministic, statistical otherwise."""
    b = brian
    b.device.reinit()
    b.device.activate()
    b.prefs["devices.b200.construction"] = "sharded"
    try:
        b.set_device("b200", directory=project_dir, build_on_run=False)
        b.prefs.codegen.cpp.extra_compile_args_gcc = list(models.STRICT_GCC_FLAGS)
        b.seed(3)
        G = b.NeuronGroup(400, "v : 1", threshold="v > 1", reset="v = 0", name="sc_G")
        H = b.NeuronGroup(300, "v : 1", name="sc_H")
        S1 = b.Synapses(G, H, "w : 1", on_pre="v_post += w", name="sc_dense")
        S1.connect(p=0.4)
        S1.w = "rand()"
        S2 = b.Synapses(G[100:300], H[50:250], on_pre="v_post += 1", name="sc_cond")
        S2.connect(condition="i != j", p=0.1)
        S3 = b.Synapses(G, H, on_pre="v_post += 1", name="sc_range")
        S3.connect(j="k for k in range(i % 7, N_post, 7)")
        S4 = b.Synapses(G, G, on_pre="v_post += 1", name="sc_one")
        S4.connect(j="i")
        net = b.Network(G, H, S1, S2, S3, S4)
        net.run(1 * b.ms, namespace={})
        b.device.build(directory=project_dir, compile=True, run=True, with_output=False)
        i1, j1 = np.asarray(S1.i[:]), np.asarray(S1.j[:])
        assert _binomial_ok(len(i1), 400 * 300, 0.4)
        assert np.all(np.diff(i1.astype(np.int64) * 300 + j1) > 0)
        w = np.asarray(S1.w[:])
        assert 0.0 <= w.min() and w.max() < 1.0 and abs(w.mean() - 0.5) < 5 / np.sqrt(12 * len(w))
        i2, j2 = np.asarray(S2.i[:]), np.asarray(S2.j[:])
        assert np.all(i2 != j2) and i2.max() < 200 and j2.max() < 200
        assert _binomial_ok(len(i2), 200 * 200 - 200, 0.1)
        i3, j3 = np.asarray(S3.i[:]), np.asarray(S3.j[:])
        want = [(i, k) for i in range(400) for k in range(i % 7, 300, 7)]
        assert np.array_equal(i3, [p[0] for p in want]) and np.array_equal(j3, [p[1] for p in want])
        assert np.array_equal(np.asarray(S4.i[:]), np.arange(400))
        assert np.array_equal(np.asarray(S4.j[:]), np.arange(400))
    finally:
        b.prefs["devices.b200.construction"] = "reference"


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_world_independence():
    """Brunel with heterogeneous delays, built per rank on 2 GPUs, against the same script on 1
    GPU: identical synapses (indices, delays), identical spike trains, bit-identical state."""
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(__file__), "run_multigpu_case.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29519", script, "--sharded", "brunel_hetero",
           "synapses_only_short"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert "MULTIGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
