"""ctypes front of ``oracle/hotpath_oracle.c`` -- the CPU restatement of the reference's
per-timestep hot loop.  TEST INFRASTRUCTURE ONLY (used by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` leg; the product under ``brian2_b200/`` never imports it).

Parity status: pinned, see the header of hotpath_oracle.c and tests/test_oracle.py.

The propagator coefficients are computed exactly as the reference's generated scalar code does
(``_lio_k`` lines hoisted by ``codegen/optimisation.py``; the form of the expressions is what
``stateupdaters/exact.py:173`` produces for these equations).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hotpath_oracle.c")
LIB = os.path.join(HERE, "_build", "libhotpath_oracle.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11",
                               SRC, "-o", LIB, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oq_create.restype = ctypes.c_void_p
        _lib.oq_create.argtypes = [_dp, ctypes.c_int, _ip, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        _lib.oq_destroy.argtypes = [ctypes.c_void_p]
        _lib.oq_push.argtypes = [ctypes.c_void_p, _ip, ctypes.c_int]
        _lib.oq_peek.restype = _ip
        _lib.oq_peek.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        _lib.oq_advance.argtypes = [ctypes.c_void_p]
        _lib.oracle_lif_run_flat.restype = ctypes.c_longlong
        _lib.oracle_stdp_run.restype = ctypes.c_longlong
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class SpikeQueue:
    """``CSpikeQueue`` restated (synapses/spikequeue.h:14-206) with the interface of the
    reference's Python wrapper (synapses/cythonspikequeue.pyx): prepare / push / peek / advance."""

    def __init__(self, source_start, source_end):
        self.source_start, self.source_end = int(source_start), int(source_end)
        self._q = None

    def prepare(self, delays, dt, synapse_sources):
        delays = np.ascontiguousarray(np.atleast_1d(delays), dtype=np.float64)
        sources = np.ascontiguousarray(synapse_sources, dtype=np.int32)
        if self._q is not None:
            lib().oq_destroy(self._q)
        self._q = lib().oq_create(_d(delays), len(delays), _i(sources), len(sources), float(dt),
                                  self.source_start, self.source_end)

    def push(self, spikes):
        spikes = np.ascontiguousarray(spikes, dtype=np.int32)
        lib().oq_push(self._q, _i(spikes), len(spikes))

    def peek(self):
        n = ctypes.c_int(0)
        p = lib().oq_peek(self._q, ctypes.byref(n))
        return np.array([p[k] for k in range(n.value)], dtype=np.int32)

    def advance(self):
        lib().oq_advance(self._q)

    def __del__(self):
        if self._q is not None and _lib is not None:
            _lib.oq_destroy(self._q)
            self._q = None


def _timestep(t, dt):
    return int((t + 1e-3 * dt) / dt)


def cuba_coefficients(dt, taum=0.02, taue=0.005, taui=0.01, El=-0.049, refractory=0.005):
    """The 15 ``_lio_k`` of the CUBA state updater (method='exact'), SURVEY.md App. A.2."""
    exp = math.exp
    l2 = exp(1.0 * (-dt) / taue)
    l3 = exp(1.0 * (-dt) / taui)
    l4 = taue * exp(1.0 * (-dt) / taue)
    l5 = -exp(1.0 * dt / taue)
    l6 = 1.0 * dt / taum
    l7 = 1.0 * (-dt) / taum
    l8 = 0.0 - taum
    l9 = taui * exp(1.0 * (-dt) / taui)
    l10 = -exp(1.0 * dt / taui)
    l11 = El - El
    l12 = El - (El * exp(l7))
    l13 = 1.0 * ((l4 * (l5 + exp(l6))) * exp(l7)) / (l8 + taue)
    l14 = 1.0 * ((l9 * (l10 + exp(l6))) * exp(l7)) / (l8 + taui)
    l15 = exp(l7)
    return dict(ref_steps=_timestep(refractory, dt), a_ge=l2, a_gi=l3, c_ref=l11, c0=l12, c_ge=l13,
                c_gi=l14, c_v=l15)


def brunel_coefficients(dt, tau=0.02, mu_ext=0.04, refractory=0.002):
    """``_lio_1.._lio_5`` of the deterministic Brunel updater ``dv/dt = (-v + mu_ext)/tau``."""
    l2 = 1.0 * (-dt) / tau
    return dict(ref_steps=_timestep(refractory, dt), a_ge=0.0, a_gi=0.0, c_ref=mu_ext - mu_ext,
                c0=mu_ext - (mu_ext * math.exp(l2)), c_ge=0.0, c_gi=0.0, c_v=math.exp(l2))


def lif_run(v, ge, gi, coeffs, v_thresh, v_reset, pathways, dt, n_steps, lastspike=None,
            not_refractory=None, cap_spikes=None):
    """Run the LIF-family network.  ``pathways``: list of dicts with keys pre, post, delay
    (array of 1 or n_syn seconds), source (start, stop), target ('v'|'ge'|'gi'), weight (float)
    or w (array).  Returns a dict with the final state and the monitors."""
    N = len(v)
    use_ge_gi = ge is not None
    v = np.array(v, dtype=np.float64)
    ge = np.array(ge if use_ge_gi else np.zeros(N), dtype=np.float64)
    gi = np.array(gi if use_ge_gi else np.zeros(N), dtype=np.float64)
    lastspike = np.array(lastspike if lastspike is not None else np.full(N, -1e4), dtype=np.float64)
    not_refractory = np.array(not_refractory if not_refractory is not None else np.ones(N), dtype=np.int8)
    mpar = np.array([coeffs["ref_steps"], coeffs["a_ge"], coeffs["a_gi"], coeffs["c_ref"], coeffs["c0"],
                     coeffs["c_ge"], coeffs["c_gi"], coeffs["c_v"], v_thresh, v_reset], dtype=np.float64)
    P = len(pathways)
    keep = []
    PP = ctypes.c_void_p * max(P, 1)
    pre, post, delay, w = PP(), PP(), PP(), PP()
    n_delays = (ctypes.c_int * max(P, 1))()
    n_syn = (ctypes.c_int * max(P, 1))()
    s0 = (ctypes.c_int * max(P, 1))()
    s1 = (ctypes.c_int * max(P, 1))()
    tv = (ctypes.c_int * max(P, 1))()
    weight = (ctypes.c_double * max(P, 1))()
    for k, pw in enumerate(pathways):
        a_pre = np.ascontiguousarray(pw["pre"], dtype=np.int32)
        a_post = np.ascontiguousarray(pw["post"], dtype=np.int32)
        a_delay = np.ascontiguousarray(np.atleast_1d(pw.get("delay", 0.0)), dtype=np.float64)
        keep += [a_pre, a_post, a_delay]
        pre[k], post[k], delay[k] = a_pre.ctypes.data, a_post.ctypes.data, a_delay.ctypes.data
        n_delays[k], n_syn[k] = len(a_delay), len(a_pre)
        s0[k], s1[k] = pw["source"]
        tv[k] = {"v": 0, "ge": 1, "gi": 2}[pw["target"]]
        weight[k] = float(pw.get("weight", 0.0))
        if pw.get("w") is not None:
            a_w = np.ascontiguousarray(pw["w"], dtype=np.float64)
            keep.append(a_w)
            w[k] = a_w.ctypes.data
        else:
            w[k] = None
    cap = int(cap_spikes if cap_spikes is not None else max(1024, N * n_steps // 20))
    while True:
        vv, gge, ggi, ls, nr = v.copy(), ge.copy(), gi.copy(), lastspike.copy(), not_refractory.copy()
        mon_i = np.zeros(cap, dtype=np.int32)
        mon_t = np.zeros(cap, dtype=np.float64)
        count = np.zeros(N, dtype=np.int32)
        rate = np.zeros(n_steps, dtype=np.float64)
        events = ctypes.c_double(0.0)
        n = lib().oracle_lif_run_flat(
            ctypes.c_int(N), _d(vv), _d(gge), _d(ggi), _d(ls), nr.ctypes.data_as(ctypes.c_char_p), _d(mpar),
            ctypes.c_int(1 if use_ge_gi else 0), ctypes.c_int(P), pre, post, delay, n_delays, n_syn, s0, s1, tv,
            weight, w, ctypes.c_double(dt), ctypes.c_longlong(n_steps), _i(mon_i), _d(mon_t),
            ctypes.c_longlong(cap), _i(count), _d(rate), ctypes.byref(events))
        if n <= cap:
            break
        cap = int(n)
    return dict(v=vv, ge=gge, gi=ggi, lastspike=ls, not_refractory=nr, spikes_i=mon_i[:n], spikes_t=mon_t[:n],
                spikes_count=count, rate=rate, events=events.value)


def stdp_run(x, rate, w, par, dt, n_steps, v0):
    """Song-Abbott STDP network of tests/models.py:stdp.  ``par``: dict of the namespace."""
    N = len(x)
    x = np.array(x, dtype=np.float64)
    rate = np.ascontiguousarray(rate, dtype=np.float64)
    w = np.array(w, dtype=np.float64)
    Apre, Apost, lastupdate = np.zeros(N), np.zeros(N), np.zeros(N)
    v1, ge1 = np.array([v0], dtype=np.float64), np.zeros(1)
    p = np.array([par[k] for k in ("taue", "taum", "El", "Ee", "vt", "vr", "taupre", "taupost", "dApre",
                                   "dApost", "gmax")], dtype=np.float64)
    cap_in = max(1024, int(N * n_steps * dt * 30) + 1024)
    in_i, in_t = np.zeros(cap_in, dtype=np.int32), np.zeros(cap_in)
    out_t = np.zeros(n_steps)
    n_out = ctypes.c_longlong(0)
    n_in = lib().oracle_stdp_run(ctypes.c_int(N), _d(x), _d(rate), _d(v1), _d(ge1), _d(w), _d(Apre), _d(Apost),
                                 _d(lastupdate), _d(p), ctypes.c_double(dt), ctypes.c_longlong(n_steps),
                                 _i(in_i), _d(in_t), ctypes.c_longlong(cap_in), _d(out_t),
                                 ctypes.c_longlong(n_steps), ctypes.byref(n_out))
    assert n_in <= cap_in
    return dict(w=w, Apre=Apre, Apost=Apost, lastupdate=lastupdate, v=v1, ge=ge1, x=x,
                in_spikes_i=in_i[:n_in], in_spikes_t=in_t[:n_in], spikes_t=out_t[:n_out.value])


def hh_run(state, par, ce, ci, dt, n_steps, Ne):
    """COBAHH network of tests/models.py:cobahh.  ``state``: dict v, ge, gi, m, n, h (initial);
    ``par``: dict of the namespace (floats in SI units) + 'refractory'; ``ce``/``ci``: (pre, post)
    absolute indices.  Returns the final state and the spike monitor."""
    N = len(state["v"])
    s = {k: np.array(state[k], dtype=np.float64) for k in ("v", "ge", "gi", "m", "n", "h")}
    lastspike = np.full(N, -1e4)
    not_refractory = np.ones(N, dtype=np.int8)
    p = np.array([par[k] for k in ("Cm", "gl", "El", "EK", "ENa", "g_na", "g_kd", "VT", "taue", "taui",
                                   "Ee", "Ei", "we", "wi", "refractory")], dtype=np.float64)
    ce_pre, ce_post = (np.ascontiguousarray(a, dtype=np.int32) for a in ce)
    ci_pre, ci_post = (np.ascontiguousarray(a, dtype=np.int32) for a in ci)
    cap = max(4096, N * n_steps // 50)
    mon_i, mon_t = np.zeros(cap, dtype=np.int32), np.zeros(cap)
    count = np.zeros(N, dtype=np.int32)
    events = ctypes.c_double(0.0)
    L = lib()
    L.oracle_hh_run.restype = ctypes.c_longlong
    nrec = L.oracle_hh_run(
        ctypes.c_int(N), ctypes.c_int(Ne), _d(s["v"]), _d(s["ge"]), _d(s["gi"]), _d(s["m"]), _d(s["n"]),
        _d(s["h"]), _d(lastspike), not_refractory.ctypes.data_as(ctypes.c_char_p), _d(p),
        _i(ce_pre), _i(ce_post), ctypes.c_int(len(ce_pre)), _i(ci_pre), _i(ci_post), ctypes.c_int(len(ci_pre)),
        ctypes.c_double(dt), ctypes.c_longlong(n_steps), _i(mon_i), _d(mon_t), ctypes.c_longlong(cap),
        _i(count), ctypes.byref(events))
    assert nrec <= cap
    out = dict(s)
    out.update(spikes_i=mon_i[:nrec], spikes_t=mon_t[:nrec], spikes_count=count, events=events.value)
    return out
