"""CPU checks of bench.py's host-side helpers (no GPU, no compiled code: Brian2's numpy runtime)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_event_count_of_the_reference_arm(brian):
    """`_outdegree_events` (the numerator of the reference arm's events/s) against a brute-force
    count: every spike of a pathway's event source times the synapses attached to that neuron,
    over on_pre AND on_post pathways, subgroup sources included, spikes before `t_from` excluded."""
    import bench

    b = brian
    b.device.reinit()
    b.device.activate()
    b.set_device("runtime")
    b.prefs.codegen.target = "numpy"
    b.defaultclock.dt = 0.1 * b.ms
    b.seed(5)
    G = b.NeuronGroup(60, "dv/dt = (1.5 - v)/(3*ms) : 1", threshold="v > 1", reset="v = 0", method="exact")
    G.v = "rand()"
    H = b.NeuronGroup(20, "dv/dt = (1.2 - v)/(5*ms) : 1\nx : 1", threshold="v > 1", reset="v = 0", method="exact")
    S = b.Synapses(G[10:50], H, "w : 1", on_pre="x_post += 1", on_post="w += 1")
    S.connect(p=0.3)
    mg, mh = b.SpikeMonitor(G), b.SpikeMonitor(H)
    net = b.Network(G, H, S, mg, mh)
    net.run(20 * b.ms)
    objs = dict(G=G, H=H, S=S, spikes=mg, post_spikes=mh)
    t_from = 8e-3
    events, nspikes = bench._outdegree_events(b, objs, t_from)
    pre = np.asarray(S.i[:]) + 10          # absolute index in G
    post = np.asarray(S.j[:])
    brute = 0
    for i, t in zip(np.asarray(mg.i[:]), np.asarray(mg.t_[:])):
        if t >= t_from - 1e-12:
            brute += int(np.sum(pre == i))
    for i, t in zip(np.asarray(mh.i[:]), np.asarray(mh.t_[:])):
        if t >= t_from - 1e-12:
            brute += int(np.sum(post == i))
    assert brute > 0 and events == brute
    b.device.reinit()
    b.device.activate()


def test_weak_scaling_keeps_synapses_per_neuron():
    import bench

    k = bench._scaled(dict(N=256000, p=80.0 / 256000), 8)
    assert k["N"] == 8 * 256000 and abs(k["N"] * k["p"] - 80.0) < 1e-9
    k = bench._scaled(dict(N_E=100000, epsilon=0.008, deterministic=True), 8)
    assert k["N_E"] == 800000 and abs(k["N_E"] * k["epsilon"] - 800.0) < 1e-9
    assert bench._scaled(dict(N=10, p=0.5), 1) == dict(N=10, p=0.5)


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    import bench

    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        per_step = json.load(f)["cobahh_256k"]["dram_bytes_per_timestep"]
    assert bench._ncu_traffic("cobahh_256k", 4000) == per_step * 4000
    assert bench._ncu_traffic("no_such_workload", 4000) is None
    # every workload names its model and its algorithmic bytes
    import models

    for name, (model, kwds, b_neuron, b_event) in bench.WORKLOADS.items():
        assert model in models.MODELS, name
        assert b_event in (20.0, 28.0, 84.0) and b_neuron >= 0.0, name
