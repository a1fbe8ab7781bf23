"""The CPU restatement (oracle/hotpath_oracle.c) against the reference: its own known-answer
tests for the spike queue and delayed delivery, and fixtures produced by the unmodified
reference's cpp_standalone device (tests/golden/make_oracle_fixtures.py).  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hotpath_oracle as ho  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DT = 1e-4


def test_spikequeue_kat_one_to_one_and_all_to_all():
    """brian2/tests/test_spikequeue.py:36-61, same inputs and expectations."""
    N = 100
    data = np.arange(N, dtype=np.int32)
    q = ho.SpikeQueue(0, N)
    q.prepare(data * DT, DT, data)
    q.push(np.arange(N, dtype=np.int32))
    for i in range(N):
        assert np.array_equal(q.peek(), [i])
        q.advance()
    for i in range(N):
        assert len(q.peek()) == 0
        q.advance()
    data = np.repeat(np.arange(N, dtype=np.int32), N)
    q = ho.SpikeQueue(0, N)
    q.prepare(data * DT, DT, data)
    q.push(np.arange(N * N, dtype=np.int32))
    for i in range(N):
        assert np.array_equal(q.peek(), i * N + np.arange(N))
        q.advance()
    for i in range(N):
        assert len(q.peek()) == 0
        q.advance()


def test_spikequeue_against_reference_cython_queue():
    """Random pushes: the restatement and the reference's compiled CSpikeQueue (through its
    Cython wrapper, synapses/cythonspikequeue.pyx) must deliver identical buckets, in order."""
    try:
        import brian2_b200  # noqa: F401  (puts baseline/_ref on sys.path)
        from brian2.synapses.cythonspikequeue import SpikeQueue as RefQueue
    except ImportError:
        pytest.skip("reference spike queue extension not built")
    rng = np.random.RandomState(3)
    for hetero in (True, False):
        n_src, n_syn, start = 50, 700, 20
        sources = np.sort(rng.randint(start, start + n_src, n_syn)).astype(np.int32)
        delays = (rng.randint(0, 12, n_syn) * DT) if hetero else np.array([3 * DT])
        ref = RefQueue(source_start=start, source_end=start + n_src)
        ref.prepare(np.asarray(delays, dtype=np.float64), DT, sources)
        mine = ho.SpikeQueue(start, start + n_src)
        mine.prepare(delays, DT, sources)
        for step in range(60):
            ref.advance()
            mine.advance()
            spikes = np.sort(rng.choice(np.arange(0, 100), size=rng.randint(0, 15), replace=False)).astype(np.int32)
            ref.push(spikes)
            mine.push(spikes)
            assert np.array_equal(np.asarray(ref.peek()), mine.peek()), (hetero, step)


def test_all_to_one_heterogeneous_delays_kat():
    """brian2/tests/test_synapses.py:1154-1176: expected v = 3, 12, 33, 48."""
    idx = [0, 1, 4, 5, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5]
    times = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2]
    w = np.array([1, 2, 3, 4, 5, 6], dtype=float)
    delay = np.array([0, 0, 0, 1, 2, 1]) * DT
    q = ho.SpikeQueue(0, 6)
    q.prepare(delay, DT, np.arange(6, dtype=np.int32))
    v, seen = 0.0, []
    for step in range(4):
        q.advance()
        q.push(np.sort(np.array([i for i, t in zip(idx, times) if t == step], dtype=np.int32)))
        for k in q.peek():
            v += w[k]
        seen.append(v)
    assert seen == [3, 12, 33, 48]


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_cuba_fixture_bit_exact():
    g = _load("oracle_cuba_400")
    n_steps = int(round(float(g["duration"][0]) / DT))
    N = len(g["in_P_v"])
    paths = [
        dict(pre=g["in_Ce_pre"], post=g["in_Ce_post"], delay=[0.0], source=(0, N), target="ge",
             weight=(60 * 0.27 / 10) * 1e-3),
        dict(pre=g["in_Ci_pre"], post=g["in_Ci_post"], delay=[0.0], source=(0, N), target="gi",
             weight=(-20 * 4.5 / 10) * 1e-3),
    ]
    r = ho.lif_run(g["in_P_v"], g["in_P_ge"], g["in_P_gi"], ho.cuba_coefficients(DT), -0.05, -0.06,
                   paths, DT, n_steps)
    assert len(g["spikes_i"]) > 100
    assert np.array_equal(r["spikes_i"], g["spikes_i"])
    assert np.array_equal(r["spikes_t"], g["spikes_t"])
    assert np.array_equal(r["spikes_count"], g["spikes_count"])
    for k in ("v", "ge", "gi"):
        assert np.array_equal(r[k], g["P_" + k]), k


def test_brunel_hetero_delay_fixture_bit_exact():
    g = _load("oracle_brunel_500")
    n_steps = int(round(float(g["duration"][0]) / DT))
    N = len(g["in_neurons_v"])
    N_E = 400
    C_E = 40
    mV = ms = 0.001   # Brian's unit arithmetic: 20*mV is the float product 20*0.001
    J, theta, tau, gg, V_r = 0.1 * mV, 20 * mV, 20 * ms, 5.0, 10 * mV
    nu_thr = theta / (J * C_E * tau)
    nu_ext = 2.0 * nu_thr
    mu_ext = J * C_E * nu_ext * tau
    # pathways in the reference's schedule order: (when, order, name) -> "brunel_exc_pre" first
    paths = [
        dict(pre=g["in_exc_pre"], post=g["in_exc_post"], delay=g["in_exc_delay"], source=(0, N_E), target="v", weight=J),
        dict(pre=g["in_inh_pre"], post=g["in_inh_post"], delay=g["in_inh_delay"], source=(N_E, N), target="v",
             weight=-gg * J),
    ]
    r = ho.lif_run(g["in_neurons_v"], None, None, ho.brunel_coefficients(DT, tau=tau, mu_ext=mu_ext), theta, V_r,
                   paths, DT, n_steps)
    assert len(g["spikes_i"]) > 100
    assert np.array_equal(r["spikes_i"], g["spikes_i"])
    assert np.array_equal(r["spikes_t"], g["spikes_t"])
    assert np.array_equal(r["v"], g["neurons_v"])
    assert np.array_equal(r["rate"], g["rate_rate"])


def test_stdp_fixture_bit_exact():
    g = _load("oracle_stdp_200")
    n_steps = int(round(float(g["duration"][0]) / DT))
    gmax = .01
    par = dict(taue=5e-3, taum=10e-3, El=-74e-3, Ee=0.0, vt=-54e-3, vr=-60e-3, taupre=20e-3, taupost=20e-3,
               dApre=.01 * gmax, dApost=-.01 * 20e-3 / 20e-3 * 1.05 * gmax, gmax=gmax)
    r = ho.stdp_run(g["in_inp_x"], g["in_inp_rate"], g["in_S_w"], par, DT, n_steps, v0=float(g["in_neurons_v"][0]))
    assert np.array_equal(r["in_spikes_i"], g["in_spikes_i"])
    assert np.array_equal(r["in_spikes_t"], g["in_spikes_t"])
    assert np.array_equal(r["spikes_t"], g["spikes_t"])
    assert np.array_equal(r["w"], g["S_w"])
    assert np.array_equal(r["v"], g["neurons_v"])
    assert np.array_equal(r["ge"], g["neurons_ge"])


def test_cobahh_fixture_bit_exact():
    """BASELINE configs[1] family: the Hodgkin-Huxley network (exponential Euler, exprel rate
    functions, refractory threshold without reset, two delay-free conductance pathways) restated
    in C against the unmodified reference's cpp_standalone run of the same inputs: spike train
    identical and -- same libm, same grouping of the terms -- state bit-identical."""
    g = _load("oracle_cobahh_300")
    n_steps = int(round(float(g["duration"][0]) / DT))
    N = len(g["in_P_v"])
    ms = mV = 0.001
    um, cm, uF, siemens, msiemens, nS = 1e-6, 0.01, 1e-6, 1.0, 0.001, 1e-9
    area = 20000 * um ** 2
    par = dict(Cm=(1 * uF * cm ** -2) * area, gl=(5e-5 * siemens * cm ** -2) * area, El=-60 * mV, EK=-90 * mV,
               ENa=50 * mV, g_na=(100 * msiemens * cm ** -2) * area, g_kd=(30 * msiemens * cm ** -2) * area,
               VT=-63 * mV, taue=5 * ms, taui=10 * ms, Ee=0 * mV, Ei=-80 * mV, we=6 * nS, wi=67 * nS,
               refractory=3 * ms)
    state = {k: g["in_P_" + k] for k in ("v", "ge", "gi", "m", "n", "h")}
    r = ho.hh_run(state, par, (g["in_Ce_pre"], g["in_Ce_post"]), (g["in_Ci_pre"], g["in_Ci_post"]), DT,
                  n_steps, Ne=int(0.8 * N))
    assert len(g["spikes_i"]) > 100
    assert np.array_equal(r["spikes_i"], g["spikes_i"])
    assert np.array_equal(r["spikes_t"], g["spikes_t"])
    assert np.array_equal(r["spikes_count"], g["spikes_count"])
    for k in ("v", "ge", "gi", "m", "n", "h"):
        np.testing.assert_allclose(r[k], g["P_" + k], rtol=1e-12, atol=0, err_msg=k)
        assert np.array_equal(r[k], g["P_" + k]), k
