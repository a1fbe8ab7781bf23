"""Brian2 model scripts shared by the golden-vector generator (reference ``cpp_standalone``), the
parity tests (``b200`` device) and ``bench.py``.  Every builder takes the already-imported
``brian2`` namespace module ``b`` and returns a dict of the objects whose state is compared.

Sources of the models (reference repository): ``examples/CUBA.py``, ``examples/COBAHH.py``,
``examples/frompapers/Brunel_2000.py``, ``examples/synapses/STDP.py``,
``brian2/tests/features/speed.py:263-326`` (SynapsesOnly stress).
"""
import numpy as np

#: strict floating-point flags for the parity oracle
STRICT_GCC_FLAGS = ["-w", "-O3", "-std=c++17", "-ffp-contract=off"]


def cuba(b, N=4000, p=0.02, duration=0.2, seed=1234, monitor=True):
    """examples/CUBA.py (4000 LIF, 2 % connectivity, exact integration)."""
    b.seed(seed)
    ms, mV = b.ms, b.mV
    taum = 20 * ms; taue = 5 * ms; taui = 10 * ms  # noqa: E702
    Vt = -50 * mV; Vr = -60 * mV; El = -49 * mV  # noqa: E702
    eqs = """
    dv/dt  = (ge+gi-(v-El))/taum : volt (unless refractory)
    dge/dt = -ge/taue : volt
    dgi/dt = -gi/taui : volt
    """
    P = b.NeuronGroup(N, eqs, threshold="v>Vt", reset="v = Vr", refractory=5 * ms, method="exact",
                      name="cuba_P", namespace=dict(taum=taum, taue=taue, taui=taui, Vt=Vt, Vr=Vr, El=El))
    P.v = "Vr + rand() * (Vt - Vr)"
    P.ge = 0 * mV
    P.gi = 0 * mV
    we = (60 * 0.27 / 10) * mV
    wi = (-20 * 4.5 / 10) * mV
    Ne = int(0.8 * N)
    Ce = b.Synapses(P, P, on_pre="ge += we", namespace=dict(we=we), name="cuba_Ce")
    Ci = b.Synapses(P, P, on_pre="gi += wi", namespace=dict(wi=wi), name="cuba_Ci")
    Ce.connect(f"i<{Ne}", p=p)
    Ci.connect(f"i>={Ne}", p=p)
    objs = dict(P=P, Ce=Ce, Ci=Ci)
    if monitor:
        objs["spikes"] = b.SpikeMonitor(P, name="cuba_spikes")
    net = b.Network(*objs.values())
    objs["net"] = net
    objs["duration"] = duration
    objs["state"] = [("P", "v"), ("P", "ge"), ("P", "gi")]
    return objs


def cobahh(b, N=4000, duration=0.1, seed=1234, monitor=True, n_syn_per_neuron=80.0, trace=(1, 10, 100)):
    """examples/COBAHH.py equations; connectivity p = 80/N as brian2/tests/features/speed.py:198."""
    b.seed(seed)
    ms, mV, cm, um, msiemens, uF, siemens, nS = b.ms, b.mV, b.cm, b.um, b.msiemens, b.uF, b.siemens, b.nS
    area = 20000 * um ** 2
    ns = dict(
        Cm=(1 * uF * cm ** -2) * area, gl=(5e-5 * siemens * cm ** -2) * area, El=-60 * mV,
        EK=-90 * mV, ENa=50 * mV, g_na=(100 * msiemens * cm ** -2) * area,
        g_kd=(30 * msiemens * cm ** -2) * area, VT=-63 * mV, taue=5 * ms, taui=10 * ms,
        Ee=0 * mV, Ei=-80 * mV, we=6 * nS, wi=67 * nS,
    )
    eqs = b.Equations("""
    dv/dt = (gl*(El-v)+ge*(Ee-v)+gi*(Ei-v)-
             g_na*(m*m*m)*h*(v-ENa)-
             g_kd*(n*n*n*n)*(v-EK))/Cm : volt
    dm/dt = alpha_m*(1-m)-beta_m*m : 1
    dn/dt = alpha_n*(1-n)-beta_n*n : 1
    dh/dt = alpha_h*(1-h)-beta_h*h : 1
    dge/dt = -ge*(1./taue) : siemens
    dgi/dt = -gi*(1./taui) : siemens
    alpha_m = 0.32*(mV**-1)*4*mV/exprel((13*mV-v+VT)/(4*mV))/ms : Hz
    beta_m = 0.28*(mV**-1)*5*mV/exprel((v-VT-40*mV)/(5*mV))/ms : Hz
    alpha_h = 0.128*exp((17*mV-v+VT)/(18*mV))/ms : Hz
    beta_h = 4./(1+exp((40*mV-v+VT)/(5*mV)))/ms : Hz
    alpha_n = 0.032*(mV**-1)*5*mV/exprel((15*mV-v+VT)/(5*mV))/ms : Hz
    beta_n = .5*exp((10*mV-v+VT)/(40*mV))/ms : Hz
    """)
    P = b.NeuronGroup(N, model=eqs, threshold="v>-20*mV", refractory=3 * ms,
                      method="exponential_euler", namespace=ns, name="hh_P")
    Ne = int(0.8 * N)
    Pe = P[:Ne]
    Pi = P[Ne:]
    Ce = b.Synapses(Pe, P, on_pre="ge+=we", namespace=ns, name="hh_Ce")
    Ci = b.Synapses(Pi, P, on_pre="gi+=wi", namespace=ns, name="hh_Ci")
    Ce.connect(p=n_syn_per_neuron / N)
    Ci.connect(p=n_syn_per_neuron / N)
    P.v = "El + (randn() * 5 - 5)*mV"
    P.ge = "(randn() * 1.5 + 4) * 10.*nS"
    P.gi = "(randn() * 12 + 20) * 10.*nS"
    objs = dict(P=P, Ce=Ce, Ci=Ci)
    if monitor:
        objs["spikes"] = b.SpikeMonitor(P, name="hh_spikes")
        if trace:
            objs["trace"] = b.StateMonitor(P, "v", record=[t for t in trace if t < N], name="hh_trace")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("P", "v"), ("P", "ge"), ("P", "gi"), ("P", "m"), ("P", "n"), ("P", "h")]
    return objs


def brunel(b, N_E=800, gamma=0.25, epsilon=0.1, duration=0.1, seed=4321, hetero_delays=True,
           deterministic=True, monitor=True):
    """Brunel (2000) sparse E/I LIF network, examples/frompapers/Brunel_2000.py:28-80, with
    heterogeneous integer delays.  ``deterministic=True`` replaces the Poisson drive by the
    constant mean drive so that spike trains are comparable bit for bit."""
    b.seed(seed)
    ms, mV, Hz = b.ms, b.mV, b.Hz
    N_I = int(round(gamma * N_E))
    N = N_E + N_I
    C_E = int(epsilon * N_E)
    C_ext = C_E
    tau = 20 * ms; theta = 20 * mV; V_r = 10 * mV; tau_rp = 2 * ms  # noqa: E702
    J = 0.1 * mV; D = 1.5 * ms; g = 5.0; nu_ext_over_nu_thr = 2.0  # noqa: E702
    nu_thr = theta / (J * C_E * tau)
    nu_ext = nu_ext_over_nu_thr * nu_thr
    ns = dict(tau=tau, theta=theta, V_r=V_r, J=J, g=g, mu_ext=J * C_ext * nu_ext * tau)
    if deterministic:
        eqs = "dv/dt = (-v + mu_ext)/tau : volt (unless refractory)"
    else:
        eqs = "dv/dt = -v/tau : volt (unless refractory)"
    neurons = b.NeuronGroup(N, eqs, threshold="v > theta", reset="v = V_r", refractory=tau_rp,
                            method="exact", namespace=ns, name="brunel_neurons")
    neurons.v = "rand() * theta"
    exc = b.Synapses(neurons[:N_E], neurons, on_pre="v += J", namespace=ns, name="brunel_exc")
    inh = b.Synapses(neurons[N_E:], neurons, on_pre="v += -g*J", namespace=ns, name="brunel_inh")
    exc.connect(p=epsilon)
    inh.connect(p=epsilon)
    if hetero_delays:
        exc.delay = "(1 + int(rand()*20)) * 0.1*ms"
        inh.delay = "(1 + int(rand()*20)) * 0.1*ms"
    else:
        exc.delay = D
        inh.delay = D
    objs = dict(neurons=neurons, exc=exc, inh=inh)
    if not deterministic:
        objs["drive"] = b.PoissonInput(target=neurons, target_var="v", N=C_ext, rate=nu_ext, weight=J)
    if monitor:
        objs["spikes"] = b.SpikeMonitor(neurons, name="brunel_spikes")
        objs["rate"] = b.PopulationRateMonitor(neurons, name="brunel_rate")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("neurons", "v")]
    return objs


def stdp(b, N=1000, duration=0.2, seed=99, raster_seed=7, monitor=True):
    """Song-Abbott STDP (examples/synapses/STDP.py:30-62).  The Poisson input is replaced by a
    pre-drawn raster fed through a SpikeGeneratorGroup-free deterministic source: a group whose
    threshold compares against a pre-drawn per-neuron phase, so the run is deterministic."""
    b.seed(seed)
    ms, mV, Hz = b.ms, b.mV, b.Hz
    taum = 10 * ms; taupre = 20 * ms; taupost = taupre  # noqa: E702
    Ee = 0 * mV; vt = -54 * mV; vr = -60 * mV; El = -74 * mV; taue = 5 * ms  # noqa: E702
    gmax = .01
    dApre = .01
    dApost = -dApre * taupre / taupost * 1.05
    dApost *= gmax
    dApre *= gmax
    ns = dict(taum=taum, taupre=taupre, taupost=taupost, Ee=Ee, vt=vt, vr=vr, El=El, taue=taue,
              gmax=gmax, dApre=dApre, dApost=dApost)
    # deterministic regular-firing inputs with heterogeneous periods (15 Hz mean)
    rng = np.random.RandomState(raster_seed)
    inp = b.NeuronGroup(N, "dx/dt = rate : 1\nrate : Hz", threshold="x > 1", reset="x = 0",
                        method="euler", name="stdp_inputs")
    inp.rate = rng.uniform(5, 25, N) * Hz
    inp.x = rng.uniform(0, 1, N)
    neurons = b.NeuronGroup(1, """dv/dt = (ge * (Ee-v) + El - v) / taum : volt
                                  dge/dt = -ge / taue : 1""",
                            threshold="v>vt", reset="v = vr", method="euler", namespace=ns,
                            name="stdp_neurons")
    neurons.v = vr
    S = b.Synapses(inp, neurons,
                   """w : 1
                      dApre/dt = -Apre / taupre : 1 (event-driven)
                      dApost/dt = -Apost / taupost : 1 (event-driven)""",
                   on_pre="""ge += w
                             Apre += dApre
                             w = clip(w + Apost, 0, gmax)""",
                   on_post="""Apost += dApost
                              w = clip(w + Apre, 0, gmax)""", namespace=ns, name="stdp_S")
    S.connect()
    S.w = "rand() * gmax"
    objs = dict(inp=inp, neurons=neurons, S=S)
    if monitor:
        objs["spikes"] = b.SpikeMonitor(neurons, name="stdp_spikes")
        objs["in_spikes"] = b.SpikeMonitor(inp, name="stdp_in_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("S", "w"), ("neurons", "v"), ("neurons", "ge")]
    return objs


def synapses_only(b, N=20000, p=0.2, rate_hz=100.0, duration=0.01, seed=11, delay_steps=0, hetero_bins=0):
    """Propagation stress test (brian2/tests/features/speed.py:263-326 `SynapsesOnly`): M source
    neurons that spike every step, N targets, `w += 1.0` per event."""
    b.seed(seed)
    dt = float(b.defaultclock.dt)
    M = max(1, int(rate_hz * N * dt))
    G = b.NeuronGroup(M, "v:1", threshold="True", name="so_sources")
    H = b.NeuronGroup(N, "w:1", name="so_targets")
    S = b.Synapses(G, H, on_pre="w += 1.0", name="so_S")
    S.connect(True, p=p)
    if delay_steps:
        S.delay = delay_steps * b.defaultclock.dt
    if hetero_bins:     # delays 0 .. hetero_bins-1 steps, by target
        S.delay = f"(j % {int(hetero_bins)}) * 0.1*ms"
    objs = dict(G=G, H=H, S=S)
    objs["net"] = b.Network(G, H, S)
    objs["duration"] = duration
    objs["state"] = [("H", "w")]
    return objs


def spikegen(b, N=200, n_spikes=3000, duration=0.05, seed=5, raster_seed=3, period_ms=None):
    """SpikeGeneratorGroup (templates/spikegenerator.cpp) driving LIF neurons through synapses
    with heterogeneous delays; `w` is a per-synapse counter changed by on_pre and recorded with
    a StateMonitor of the Synapses.  All increments are exactly representable, so every sum is
    exact in any order."""
    b.seed(seed)
    ms = b.ms
    rng = np.random.RandomState(raster_seed)
    horizon = int(round((period_ms if period_ms else duration * 1e3) * 10))   # time bins
    # at most one spike per neuron per bin; spike times are exact bin centres
    pairs = rng.choice(N * horizon, size=min(n_spikes, N * horizon), replace=False)
    idx = (pairs % N).astype(np.int32)
    bins = (pairs // N).astype(np.int64)
    kwds = dict(period=period_ms * ms) if period_ms else {}
    SG = b.SpikeGeneratorGroup(N, idx, bins * 0.1 * ms, name="sg_source", **kwds)
    G = b.NeuronGroup(N, "dv/dt = -v/(10*ms) : 1 (unless refractory)", threshold="v > 1", reset="v = 0",
                      refractory=2 * ms, method="exact", name="sg_neurons")
    S = b.Synapses(SG, G, "w : 1", on_pre="v_post += 0.25\nw += 0.125", name="sg_S")
    S.connect(p=0.05)
    S.delay = "(int(rand()*5)) * 0.1*ms"
    objs = dict(SG=SG, G=G, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="sg_spikes")
    objs["in_spikes"] = b.SpikeMonitor(SG, name="sg_in_spikes")
    objs["wtrace"] = b.StateMonitor(S, "w", record=[0, 7, 100], name="sg_wtrace")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("S", "w")]
    return objs


def gapjunction(b, N=300, p=0.1, duration=0.05, seed=21):
    """Summed variable (templates/summed_variable.cpp): electrical coupling
    `Igap_post = w*(v_pre - v_post) : 1 (summed)` between LIF neurons."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """dv/dt = (I0 - v + Igap)/(10*ms) : 1
                            I0 : 1
                            Igap : 1""", threshold="v > 1", reset="v = 0", method="euler", name="gj_neurons")
    G.v = "rand()"
    G.I0 = "0.8 + 0.6*rand()"
    S = b.Synapses(G, G, """w : 1
                            Igap_post = w*(v_pre - v_post) : 1 (summed)""", name="gj_S")
    S.connect(condition="i != j", p=p)
    S.w = "0.02*rand()"
    objs = dict(G=G, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="gj_spikes")
    objs["trace"] = b.StateMonitor(G, "Igap", record=[0, 1, N - 1], name="gj_trace")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("G", "Igap")]
    return objs


def timedarray(b, N=100, duration=0.05, seed=8):
    """2-d TimedArray stimulus (input/timedarray.py:282-345) read inside the state update."""
    b.seed(seed)
    ms = b.ms
    rng = np.random.RandomState(seed)
    stim = b.TimedArray(rng.uniform(0.5, 2.0, size=(25, 4)), dt=2 * ms, name="ta_stim")
    G = b.NeuronGroup(N, "dv/dt = (ta_stim(t, i % 4) - v)/(10*ms) : 1", threshold="v > 1", reset="v = 0",
                      method="euler", name="ta_neurons", namespace=dict(ta_stim=stim))
    G.v = "rand()"
    objs = dict(G=G)
    objs["spikes"] = b.SpikeMonitor(G, name="ta_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v")]
    return objs


def poisson_drive(b, N=2000, duration=0.1, seed=31):
    """In-loop random numbers: PoissonInput (binomial sampler) + PoissonGroup thresholder feeding LIF
    neurons.  No cross-target reproducibility (docs_sphinx/advanced/random.rst:28-39): compared
    statistically."""
    b.seed(seed)
    ms, Hz = b.ms, b.Hz
    G = b.NeuronGroup(N, "dv/dt = -v/(10*ms) : 1", threshold="v > 1", reset="v = 0", method="exact",
                      name="pd_neurons")
    PI = b.PoissonInput(G, "v", N=200, rate=20 * Hz, weight=0.03)
    PG = b.PoissonGroup(500, rates=40 * Hz, name="pd_group")
    S = b.Synapses(PG, G, on_pre="v += 0.05", name="pd_S")
    S.connect(p=0.02)
    objs = dict(G=G, PI=PI, PG=PG, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="pd_spikes")
    objs["in_spikes"] = b.SpikeMonitor(PG, name="pd_in_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v")]
    return objs


def ragged(b, N=600, duration=0.03, seed=17):
    """Stress of the propagation kernel's work distribution (templates/synapses.cu): CSR rows of
    very different lengths (0 ... N/2, `j <= i`), 45 distinct delays (two groups of 32 delay bins,
    delay 0 included), subgroup sources AND targets (absolute indices with offsets,
    synapses_create_generator.cpp:38,184), per-synapse weights, on_pre and on_post pathways.
    All increments are multiples of 1/8, so every sum is exact in any order."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """dv/dt = rate : 1
                            rate : Hz
                            x : 1
                            y : 1""", threshold="v >= 1", reset="v = 0", method="euler", name="rg_neurons")
    G.rate = "(50 + (i * 37) % 400) * Hz"
    G.v = "((i * 13) % 16) / 16.0"
    src = G[N // 4: 3 * N // 4]
    tgt = G[N // 8: 5 * N // 8]
    S = b.Synapses(src, tgt, "w : 1", on_pre="x_post += w", on_post="y_pre += 0.125\nw += 0.25",
                   name="rg_S")
    S.connect(condition="j <= i")
    S.w = "((i + 3 * j) % 8) * 0.125"
    S.pre.delay = "((i * 7 + j) % 45) * 0.1*ms"
    S.post.delay = "((i + j) % 3) * 0.1*ms"
    objs = dict(G=G, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="rg_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "x"), ("G", "y"), ("S", "w")]
    return objs


# Potjans & Diesmann (2014) cortical microcircuit: population sizes, connection probabilities
# C[target][source], external in-degrees (published parameter tables; not in the reference repo)
PD_NAMES = ["L23E", "L23I", "L4E", "L4I", "L5E", "L5I", "L6E", "L6I"]
PD_SIZES = [20683, 5834, 21915, 5479, 4850, 1065, 14395, 2948]
PD_CONN = [[0.1009, 0.1689, 0.0437, 0.0818, 0.0323, 0.0, 0.0076, 0.0],
           [0.1346, 0.1371, 0.0316, 0.0515, 0.0755, 0.0, 0.0042, 0.0],
           [0.0077, 0.0059, 0.0497, 0.1350, 0.0067, 0.0003, 0.0453, 0.0],
           [0.0691, 0.0029, 0.0794, 0.1597, 0.0033, 0.0, 0.1057, 0.0],
           [0.1004, 0.0622, 0.0505, 0.0057, 0.0831, 0.3726, 0.0204, 0.0],
           [0.0548, 0.0269, 0.0257, 0.0022, 0.0600, 0.3158, 0.0086, 0.0],
           [0.0156, 0.0066, 0.0211, 0.0166, 0.0572, 0.0197, 0.0396, 0.2252],
           [0.0364, 0.0010, 0.0034, 0.0005, 0.0277, 0.0080, 0.0658, 0.1443]]
PD_KEXT = [1600, 1500, 2100, 1900, 2000, 1900, 2900, 2100]


def potjans(b, scale=0.02, duration=0.05, seed=55, poisson=False, monitor=True):
    """Potjans-Diesmann microcircuit (BASELINE.json configs[4]): 8 populations of LIF neurons with
    exponential current synapses, fixed total number of synapses per projection
    K = log(1-C)/log(1-1/(N_pre N_post)) drawn with numpy (seeded) and handed to the reference's
    `Synapses.connect(i=..., j=...)` (synapses.py:1704-1710), normally distributed weights
    (87.8 pA +- 10 %, inhibition x -4, L4E->L23E doubled) and delays (1.5 +- 0.75 ms exc,
    0.8 +- 0.4 ms inh).  ``scale`` shrinks the population sizes (in-degrees are preserved by
    scaling K with it, as in the published down-scaling recipe without weight compensation).
    ``poisson=False`` replaces the 8 Hz Poisson background by its mean current, which makes the
    run deterministic (spike-exact comparison over short horizons); ``poisson=True`` uses
    PoissonInput (binomial sampler on the device's Philox streams: statistical comparison)."""
    b.seed(seed)
    ms, mV, pA, pF, Hz = b.ms, b.mV, b.pA, b.pF, b.Hz
    rng = np.random.RandomState(seed)
    sizes = [max(1, int(round(n * scale))) for n in PD_SIZES]
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(starts[-1])
    tau_m, tau_ref, tau_syn = 10 * ms, 2 * ms, 0.5 * ms
    C_m, v_r, v_th = 250 * pF, -65 * mV, -50 * mV
    w_ex, g, bg_rate = 87.8 * pA, 4.0, 8 * Hz
    ns = dict(tau_m=tau_m, tau_syn=tau_syn, C_m=C_m, v_r=v_r, v_th=v_th)
    eqs = """dv/dt = (v_r - v)/tau_m + (I + Iext)/C_m : volt (unless refractory)
             dI/dt = -I/tau_syn : amp
             Iext : amp (constant)"""
    G = b.NeuronGroup(N, eqs, threshold="v > v_th", reset="v = v_r", refractory=tau_ref,
                      method="exact", namespace=ns, name="pd_net")
    G.v = (-58.0 + 10.0 * rng.randn(N)) * mV
    kext = np.repeat(PD_KEXT, sizes).astype(float)
    if not poisson:   # mean of the Poisson background: K_ext * rate * w * tau_syn
        G.Iext = kext * float(bg_rate) * float(w_ex) * float(tau_syn) * b.amp
    pre_all, post_all, w_all, d_all = [], [], [], []
    for t in range(8):
        for s_ in range(8):
            c = PD_CONN[t][s_]
            if c == 0.0:
                continue
            n_pre_full, n_post_full = PD_SIZES[s_], PD_SIZES[t]
            k_full = np.log(1.0 - c) / np.log(1.0 - 1.0 / (n_pre_full * n_post_full))
            k = int(round(k_full * scale))       # in-degree preserved: K/N_post constant
            if k == 0:
                continue
            pre = rng.randint(0, sizes[s_], size=k) + starts[s_]
            post = rng.randint(0, sizes[t], size=k) + starts[t]
            exc = s_ % 2 == 0
            w_mean = float(w_ex) * (2.0 if (s_ == 2 and t == 0) else 1.0) * (1.0 if exc else -g)
            w = w_mean + 0.1 * abs(w_mean) * rng.randn(k)
            w = np.maximum(w, 0.0) if exc else np.minimum(w, 0.0)
            d = (1.5e-3 + 0.75e-3 * rng.randn(k)) if exc else (0.8e-3 + 0.4e-3 * rng.randn(k))
            d = np.maximum(np.round(d / 1e-4), 1.0) * 1e-4      # whole steps, >= dt
            pre_all.append(pre); post_all.append(post); w_all.append(w); d_all.append(d)  # noqa: E702
    pre = np.concatenate(pre_all); post = np.concatenate(post_all)  # noqa: E702
    order = np.lexsort((post, pre))              # (pre, post) order like a single connect() call
    S = b.Synapses(G, G, "w : amp", on_pre="I_post += w", name="pd_syn")
    S.connect(i=pre[order].astype(np.int32), j=post[order].astype(np.int32))
    S.w = np.concatenate(w_all)[order] * b.amp
    S.delay = np.concatenate(d_all)[order] * b.second
    objs = dict(G=G, S=S)
    if poisson:
        for p_ in range(8):
            sub = G[int(starts[p_]):int(starts[p_ + 1])]
            objs[f"bg{p_}"] = b.PoissonInput(sub, "I", N=PD_KEXT[p_], rate=bg_rate, weight=w_ex)
    if monitor:
        objs["spikes"] = b.SpikeMonitor(G, name="pd_net_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("G", "I")]
    return objs


MODELS = dict(cuba=cuba, cobahh=cobahh, brunel=brunel, stdp=stdp, synapses_only=synapses_only,
              spikegen=spikegen, gapjunction=gapjunction, timedarray=timedarray,
              poisson_drive=poisson_drive, ragged=ragged, potjans=potjans)


def submon(b, N=600, duration=0.05, seed=77):
    """Monitors of SUBGROUPS (spikemonitor.cpp:15-33 restricts the spike list to
    [_source_start, _source_stop); ratemonitor.cpp likewise; statemonitor of a subgroup) on a small
    CUBA-like network, plus a pathway whose source and target are subgroups."""
    b.seed(seed)
    ms, mV = b.ms, b.mV
    eqs = """dv/dt = (ge+gi-(v+49*mV))/(20*ms) : volt (unless refractory)
             dge/dt = -ge/(5*ms) : volt
             dgi/dt = -gi/(10*ms) : volt"""
    P = b.NeuronGroup(N, eqs, threshold="v>-50*mV", reset="v=-60*mV", refractory=5 * ms, method="exact",
                      name="sm_P")
    P.v = "-60*mV + rand()*10*mV"
    Ce = b.Synapses(P[: N * 4 // 5], P, on_pre="ge += 1.62*mV", name="sm_Ce")
    Ci = b.Synapses(P[N * 4 // 5:], P[N // 10:], on_pre="gi -= 9*mV", name="sm_Ci")
    Ce.connect(p=0.1)
    Ci.connect(p=0.1)
    objs = dict(P=P, Ce=Ce, Ci=Ci)
    objs["spikes"] = b.SpikeMonitor(P, name="sm_all")
    objs["sub_spikes"] = b.SpikeMonitor(P[N // 6: N // 2], name="sm_sub")
    objs["tail_spikes"] = b.SpikeMonitor(P[N - 50:], name="sm_tail")
    objs["sub_rate"] = b.PopulationRateMonitor(P[N // 3: 2 * N // 3], name="sm_rate")
    objs["sub_trace"] = b.StateMonitor(P[N // 2:], "v", record=[0, 5, 17], name="sm_trace")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("P", "v"), ("P", "ge"), ("P", "gi")]
    return objs


MODELS["submon"] = submon


def multiclock(b, N=200, duration=0.05, seed=41):
    """Several clocks in one network (network.cpp:134-159 `next_clocks`): neurons on the default
    clock (0.1 ms), a `run_regularly` that writes a SHARED variable every 1 ms (stateupdate.cpp:
    ALLOWS_SCALAR_WRITE), a StateMonitor every 0.5 ms, a rate monitor.  The drive only takes
    dyadic values, so every product in the update is exact and the states are bit-comparable."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """dv/dt = (drive - v)/(8*ms) : 1
                            drive : 1 (shared)
                            ticks : 1 (shared)""", threshold="v > 1", reset="v = 0", method="euler",
                      name="mc_neurons")
    G.v = "rand()"
    G.drive = 1.5
    G.run_regularly("drive = 1.25 + 0.25*(int(ticks) % 4)\nticks += 1", dt=1 * ms, name="mc_drive")
    objs = dict(G=G)
    objs["spikes"] = b.SpikeMonitor(G, name="mc_spikes")
    objs["trace"] = b.StateMonitor(G, ["v", "drive"], record=[0, 7, N - 1], dt=0.5 * ms, name="mc_trace")
    objs["rate"] = b.PopulationRateMonitor(G, name="mc_rate")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("G", "drive"), ("G", "ticks")]
    return objs


def sharedvar(b, N=300, duration=0.03, seed=43):
    """A shared variable written inside the loop on the SAME clock as everything else (persistent
    kernel: the writer is one elected thread, readers are ordered by grid barriers): a counter
    driven `run_regularly`, read by the state updater and by synaptic code."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """dv/dt = (gain - v)/(5*ms) : 1
                            gain : 1 (shared)
                            n_steps : 1 (shared)
                            x : 1""", threshold="v > 1", reset="v = 0", method="euler", name="sv_neurons")
    G.v = "rand()"
    G.gain = 1.5
    G.run_regularly("n_steps += 1\ngain = 1.25 + 0.125*(int(n_steps) % 5)", when="start", name="sv_rr")
    S = b.Synapses(G, G, on_pre="x_post += gain_pre", name="sv_S")
    S.connect(p=0.05)
    objs = dict(G=G, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="sv_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("G", "x"), ("G", "gain"), ("G", "n_steps")]
    return objs


def synstate(b, N=150, duration=0.03, seed=47):
    """Clock-driven synaptic state (stateupdate.cpp with N = number of synapses, `constant_or_scalar('N')`):
    a per-synapse conductance `g` decays every step, jumps on presynaptic spikes and drives the
    postsynaptic neuron through a summed variable (summed_variable.cpp)."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """dv/dt = (1.1 - v + I)/(10*ms) : 1
                            I : 1""", threshold="v > 1", reset="v = 0", method="euler", name="sy_neurons")
    G.v = "rand()"
    S = b.Synapses(G, G, """dg/dt = -g/(4*ms) : 1 (clock-driven)
                            I_post = 0.05*g : 1 (summed)""", on_pre="g += 1", method="euler", name="sy_S")
    S.connect(condition="i != j", p=0.1)
    S.g = "rand()"
    objs = dict(G=G, S=S)
    objs["spikes"] = b.SpikeMonitor(G, name="sy_spikes")
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", "v"), ("G", "I"), ("S", "g")]
    return objs


def poissonfn(b, N=4000, duration=0.02, seed=53):
    """`poisson(lam)` inside the loop (cpp_generator.py:661-751: multiplication sampler for lam < 10,
    PTRS for lam >= 10) -- statistical comparison only."""
    b.seed(seed)
    ms = b.ms
    G = b.NeuronGroup(N, """small : 1
                            large : 1
                            acc_small : 1
                            acc_large : 1""", name="pf_neurons")
    G.run_regularly("small = poisson(2.5)\nlarge = poisson(40.0)\nacc_small += small\nacc_large += large",
                    name="pf_rr")
    objs = dict(G=G)
    objs["net"] = b.Network(G)
    objs["duration"] = duration
    objs["state"] = [("G", "acc_small"), ("G", "acc_large"), ("G", "small"), ("G", "large")]
    return objs


def mathfuncs(b, N=2048, duration=0.01, seed=5):
    """Every libm function the hot path calls (exp, expm1, log, pow; plus tanh,
    sinh, cosh, which glibc builds from exp/expm1, and sin, cos) evaluated per
    neuron and per step in in-loop code over wide argument ranges (x in [-30, 30], p in (0, 60],
    exponents in [-6, 6]), plus the powers g++ folds at compile time (x**2, p**-1) and an integer
    power.  The arguments come from three chaotic (logistic) maps per neuron which every function
    value perturbs a little: a single result that is 1 ulp off at any step is amplified by ~2 per
    step until the whole state differs, so the final state is bit-identical to the reference's
    cpp_standalone run only if ALL evaluations (2048 neurons x 100 steps x 14 calls) were -- which
    `prefs.devices.b200.libm = 'glibc'` promises."""
    b.seed(seed)
    G = b.NeuronGroup(N, """u : 1
                            w : 1
                            s : 1
                            x : 1
                            p : 1
                            q : 1
                            y_exp : 1
                            y_expm1 : 1
                            y_rel : 1
                            y_log : 1
                            y_pow : 1
                            y_sq : 1
                            y_inv : 1
                            y_mix : 1
                            y_tanh : 1
                            y_sinh : 1
                            y_cosh : 1
                            y_sin : 1
                            y_cos : 1""", name="mf_G")
    G.u = "0.05 + 0.9*rand()"
    G.w = "0.05 + 0.9*rand()"
    G.s = "0.05 + 0.9*rand()"
    G.run_regularly("""x = -30 + 60*u
                       p = 1e-3 + 60*w**3
                       q = -6 + 12*s
                       y_exp = exp(x)
                       y_expm1 = expm1(0.1*x)
                       y_rel = exprel(0.05*x)
                       y_log = log(p)
                       y_pow = p**q
                       y_sq = x**2
                       y_inv = p**-1
                       y_mix = exp(0.01*x)**0.3 + log(1 + y_sq)
                       y_tanh = tanh(0.1*x)
                       y_sinh = sinh(0.2*x)
                       y_cosh = cosh(0.2*x)
                       y_sin = sin(x*p)
                       y_cos = cos(x*p*p)
                       u = 3.99*u*(1 - u)*(1 - 0.01*y_exp/(1 + y_exp) - 0.001*y_rel/(1 + y_rel) - 0.001*y_mix/(1 + y_mix))
                       w = 3.98*w*(1 - w)*(1 - 0.01*y_pow/(1 + y_pow) - 0.001/(1 + y_log*y_log))
                       s = 3.97*s*(1 - s)*(1 - 0.01/(1 + y_inv) - 0.001/(1 + y_expm1*y_expm1) - 0.001*y_tanh*y_tanh - 0.001/(1 + y_sinh*y_sinh) - 0.001/y_cosh - 0.001*y_sin*y_sin - 0.001*y_cos*y_cos)""", name="mf_rr")
    objs = dict(G=G)
    objs["net"] = b.Network(*objs.values())
    objs["duration"] = duration
    objs["state"] = [("G", v) for v in ("u", "w", "s", "x", "p", "q", "y_exp", "y_expm1", "y_rel", "y_log",
                                        "y_pow", "y_sq", "y_inv", "y_mix", "y_tanh", "y_sinh", "y_cosh",
                                        "y_sin", "y_cos")]
    return objs


MODELS.update(multiclock=multiclock, sharedvar=sharedvar, synstate=synstate, poissonfn=poissonfn,
              mathfuncs=mathfuncs)


def run_model(b, name, device_name, directory, build_kwds=None, prefs_update=None, n_runs=1, **model_kwds):
    """Build + run ``name`` on ``device_name``; returns (objs, results dict of numpy arrays).

    All objects carry explicit names: Brian orders code objects of the same schedule slot by
    (order, name) (core/network.py:907-909), so auto-generated names (`synapses_1`, ...) would
    make e.g. the excitatory/inhibitory delivery order -- and with it the last bits of `v` --
    depend on how many objects the process created before."""
    import gc

    gc.collect()
    b.device.reinit()
    b.device.activate()
    b.set_device(device_name, directory=directory, build_on_run=False)
    b.prefs.codegen.cpp.extra_compile_args_gcc = list(STRICT_GCC_FLAGS)
    b.prefs.devices.cpp_standalone.openmp_threads = 0
    if prefs_update:
        for k, v in prefs_update.items():
            b.prefs[k] = v
    b.defaultclock.dt = 0.1 * b.ms
    objs = MODELS[name](b, **model_kwds)
    net = objs["net"]
    for _ in range(n_runs):     # several run() calls of equal length (same total duration)
        net.run(objs["duration"] / n_runs * b.second, namespace={})
    b.device.build(directory=directory, compile=True, run=True, with_output=False, **(build_kwds or {}))
    res = collect_results(b, objs)
    return objs, res


def collect_results(b, objs):
    res = {}
    for key, obj in objs.items():
        if isinstance(obj, b.SpikeMonitor):
            res[f"{key}_i"] = np.asarray(obj.i[:]).astype(np.int32)
            res[f"{key}_t"] = np.asarray(obj.t_[:]).astype(np.float64)
            res[f"{key}_count"] = np.asarray(obj.count[:]).astype(np.int32)
        elif isinstance(obj, b.StateMonitor):
            for var in obj.record_variables:
                res[f"{key}_{var}"] = np.asarray(getattr(obj, var + "_")[:]).astype(np.float64)
            res[f"{key}_t"] = np.asarray(obj.t_[:]).astype(np.float64)
        elif isinstance(obj, b.PopulationRateMonitor):
            res[f"{key}_rate"] = np.asarray(obj.rate_[:]).astype(np.float64)
    for group, var in objs["state"]:
        res[f"{group}_{var}"] = np.asarray(getattr(objs[group], var + "_")[:]).copy()
    for key, obj in objs.items():
        if isinstance(obj, b.Synapses):
            res[f"{key}_nsyn"] = np.array([len(obj)], dtype=np.int64)
    res["last_run_time"] = np.array([b.device._last_run_time])
    return res
