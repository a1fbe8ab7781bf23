#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200 Brian2 device.

Metric (BASELINE.json): synaptic events/s (and the sim-time realtime factor) of the
per-timestep hot loop on the named synthetic networks.  Default workload = BASELINE.json
configs[1]: COBAHH (examples/COBAHH.py equations) scaled to 256k neurons, 80 synapses per neuron,
exponential_euler, fp64, SpikeMonitor + 3 voltage traces.

A bench "step" is one ``run(T)`` call of the Brian2 script, i.e. ``--sim-steps`` simulation
timesteps (dt = 0.1 ms) of the whole network: upload of all arrays (host -> device), the step
loop inside the persistent kernel, download of the written arrays.

* ``value``   events/s over the K timed steps with all inputs resident in HBM: CUDA-event time of
              the step loops only (the reference excludes its file I/O in the same way,
              templates/network.cpp:51,106-114).
* ``e2e``     the same events divided by upload + loop + download of every timed step, i.e. what a
              user of ``set_device('b200')`` observes per ``run()`` call, host buffers in and out.
* ``--impl reference``  the UNMODIFIED reference (``cpp_standalone`` + OpenMP on all host cores,
              default reference flags) on the same network, a bounded number of timesteps per step.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (model, kwargs, bytes per neuron-step, bytes per synaptic event)   [SURVEY.md 8d]
    "cobahh_256k": ("cobahh", dict(N=256000), 105.0, 20.0),
    "cobahh_4k": ("cobahh", dict(N=4000), 105.0, 20.0),
    "cuba_4k": ("cuba", dict(N=4000, p=0.02), 58.0, 20.0),
    "cuba_256k": ("cuba", dict(N=256000, p=80.0 / 256000), 58.0, 20.0),
    "brunel_100k": ("brunel", dict(N_E=80000, epsilon=0.01, deterministic=True), 33.0, 20.0),
    # BASELINE configs[2] per GPU: 125k neurons x 1000 synapses each; `--gpus 8` = 1M neurons / 1B
    # synapses, heterogeneous delays 1..20 steps, partitioned by postsynaptic neuron.
    # brunel_125k: connectivity by the reference's host connect() (bit-identical to cpp_standalone;
    # every rank builds the whole network); brunel_125k_sharded: every rank creates the synapses of
    # its own neurons on the device (prefs.devices.b200.construction = 'sharded') -- the only way
    # to the 10^9-synapse network
    "brunel_125k": ("brunel", dict(N_E=100000, epsilon=0.008, deterministic=True), 33.0, 20.0),
    "brunel_125k_sharded": ("brunel", dict(N_E=100000, epsilon=0.008, deterministic=True), 33.0, 20.0),
    "brunel_125k_poisson": ("brunel", dict(N_E=100000, epsilon=0.008, deterministic=False), 33.0, 20.0),
    # BASELINE configs[3]: Song-Abbott STDP, 100k plastic synapses onto one neuron (on_pre 84 B,
    # on_post 68 B per event; regular-firing inputs instead of Poisson so that it is deterministic)
    "stdp_100k": ("stdp", dict(N=100000), 0.0, 84.0),
    # BASELINE configs[4]: Potjans-Diesmann microcircuit, 77 169 neurons / ~3e8 synapses with
    # per-synapse weights and delays, DC background (deterministic); 50 B/neuron-step, 28 B/event
    "potjans_77k": ("potjans", dict(scale=1.0), 50.0, 28.0),
    "potjans_8k": ("potjans", dict(scale=0.1), 50.0, 28.0),
    # the same microcircuit with its Poisson background (8 PoissonInput objects, binomial sampler on
    # the device's Philox streams) instead of the mean current
    "potjans_77k_poisson": ("potjans", dict(scale=1.0, poisson=True), 50.0, 28.0),
    # propagation stress (brian2/tests/features/speed.py:263-326 SynapsesOnly): every source spikes
    # every step, `w += 1.0` per event -> the step is synaptic propagation only (20 B/event)
    "synapses_only_sparse": ("synapses_only", dict(N=100000, p=0.2, rate_hz=10.0), 0.0, 20.0),
    "synapses_only_dense": ("synapses_only", dict(N=40000, p=1.0, rate_hz=10.0), 0.0, 20.0),
    "synapses_only_highrate": ("synapses_only", dict(N=100000, p=0.2, rate_hz=100.0), 0.0, 20.0),
}

#: what actually bounds a step of each model family (ncu, profiles/): `roofline.bound` names the
#: roofline the contract asks for, this says why the fraction is what it is
LIMITERS = {
    "cobahh": "instruction issue (fp64 exp/expm1 chains, ~900 warp instructions per 32 neurons) + 2 grid barriers; "
              "working set L2 resident (DRAM 0.4 MB/step)",
    "cuba": "latency: ~14 dependent L2 round trips per step (grid size 4...296 CTAs changes nothing)",
    "brunel": "latency: delivery = chain of 4-5 dependent loads, one grid barrier per step",
    "synapses": "shared-memory pipe of the tiled delivery (LDS+STS per 32 events); index stream L2 resident",
    "potjans": "latency: delivery chain over 45 delay bins + 2 grid barriers",
    "stdp": "latency: on_pre -> barrier -> on_post over 100k synapses of one neuron",
}

#: device preferences of a workload (everything else runs with the defaults)
WORKLOAD_PREFS = {
    "brunel_125k_sharded": {"devices.b200.construction": "sharded"},
    "brunel_125k_poisson": {"devices.b200.construction": "sharded"},
}

#: extra configurations measured after the headline workload in a default run (no --workload):
#: BASELINE.json configs[2], the north-star network (N GPUs x 125k neurons / 125M synapses)
EXTRA_CONFIGS = ["brunel_125k_sharded"]

# simulation timesteps (dt = 0.1 ms) of one bench step = one run() call
DEFAULT_SIM_STEPS = {"cobahh_256k": 4000, "cuba_256k": 4000, "cuba_4k": 10000, "cobahh_4k": 10000,
                     "brunel_100k": 2000, "brunel_125k": 1000, "brunel_125k_sharded": 1000,
                     "brunel_125k_poisson": 1000, "stdp_100k": 5000, "potjans_77k": 1000, "potjans_77k_poisson": 1000,
                     "potjans_8k": 2000, "synapses_only_sparse": 1000, "synapses_only_dense": 1000,
                     "synapses_only_highrate": 500}

# timesteps per step of the reference arm (CPU): bounded so that 25 steps end within minutes
REFERENCE_SIM_STEPS = {"cobahh_256k": 500, "cuba_256k": 2000, "brunel_125k": 50, "brunel_125k_sharded": 50,
                       "brunel_100k": 50, "potjans_77k": 20, "synapses_only_highrate": 10,
                       "synapses_only_sparse": 50, "synapses_only_dense": 50}


def _n_neurons(objs):
    for key in ("P", "neurons", "H", "G"):
        if key in objs:
            return len(objs[key])
    return 0


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(workload, timesteps_per_launch):
    """DRAM bytes (read + write) of one launch of the persistent kernel, from this round's
    `ncu --set full` capture of the same workload (profiles/ncu_traffic.json: bytes per
    timestep of the captured launch x the timesteps of a bench launch); None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            per_step = json.load(f)[workload]["dram_bytes_per_timestep"]
    except (OSError, KeyError, ValueError):
        return None
    return per_step * timesteps_per_launch


class ClockSampler:
    """SM clock / throttle-reason samples (NVML, every ~2 ms) kept only when they fall inside
    the timed step loops (`windows` = [(t0, t1)] in host epoch seconds)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.samples = []
        self.index = index
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self.error = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:  # no NVML: report it, never fake a clock
            self.error = f"nvml unavailable: {ex}"
            return
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()

    def _poll(self):
        nv, h = self._nvml, self._h
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = reasons_fn(h)
                self.samples.append((time.time(), float(mhz), int(mask)))
            except Exception as ex:
                self.error = str(ex)
                return
            time.sleep(0.002)

    def stop(self, windows=None):
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "no samples"]}
        self._stop.set()
        self._thread.join()
        rows = [(mhz, mask) for (ts, mhz, mask) in self.samples
                if not windows or any(t0 <= ts <= t1 for (t0, t1) in windows)]
        where = "inside the timed step loops"
        if not rows:
            rows = [(mhz, mask) for (_, mhz, mask) in self.samples]
            where = "whole run (no sample fell inside a timed loop)"
        sm = sorted(m for m, _ in rows)
        reasons = set()
        for _, mask in rows:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": where}


def _import_brian():
    import brian2_b200  # noqa: F401
    import brian2 as b

    return b


def _build_script(b, workload, device_name, directory, sim_steps, n_runs, openmp_threads=0, strict=True,
                  scale=1):
    import models

    model, kwds, _, _ = WORKLOADS[workload]
    kwds = _scaled(kwds, scale)
    import gc

    gc.collect()
    b.device.reinit()
    b.device.activate()
    b.set_device(device_name, directory=directory, build_on_run=False)
    if strict:
        b.prefs.codegen.cpp.extra_compile_args_gcc = list(models.STRICT_GCC_FLAGS)
    else:  # the reference's own default flags (codegen/cpp_prefs.py:109-116)
        b.prefs.codegen.cpp.extra_compile_args_gcc = ["-w", "-O3", "-ffast-math", "-fno-finite-math-only",
                                                      "-march=native", "-std=c++17"]
    b.prefs.devices.cpp_standalone.openmp_threads = openmp_threads
    b.defaultclock.dt = 0.1 * b.ms
    objs = models.MODELS[model](b, duration=0.0, **kwds)
    net = objs["net"]
    T = sim_steps * 1e-4
    for r in range(n_runs):
        net.run(T * b.second, namespace={})
        if device_name == "cpp_standalone":
            b.device.insert_code(
                "main", 'std::cout << "B200BENCH_RUN " << Network::_last_run_time << std::endl;')
    return objs


def _scaled(kwds, scale):
    """Weak scaling: `scale` x the neurons, same number of synapses per neuron."""
    kwds = dict(kwds)
    if scale > 1:
        if "N" in kwds:
            if "p" in kwds:
                kwds["p"] = kwds["p"] / scale
            kwds["N"] = kwds["N"] * scale
        elif "N_E" in kwds:
            kwds["N_E"] = kwds["N_E"] * scale
            kwds["epsilon"] = kwds["epsilon"] / scale
    return kwds


def _outdegree_events(b, objs, t_from):
    """Synaptic events in [t_from, end): for every pathway (on_pre and on_post) the sum over the
    recorded spikes of its event source of the number of synapses attached to the spiking
    neuron.  Needs a SpikeMonitor on every group that drives a pathway."""
    import numpy as np

    mons = {m.source.name: m for m in objs.values() if isinstance(m, b.SpikeMonitor)}
    total, nspikes = 0.0, 0
    for obj in objs.values():
        if not isinstance(obj, b.Synapses):
            continue
        for path in obj._pathways:
            grp = path.source
            parent = getattr(grp, "source", grp)
            ends = np.asarray(obj.i[:] if path.prepost == "pre" else obj.j[:])
            if grp.name in mons:        # monitor on the (sub)group itself: relative indices
                mon, deg = mons[grp.name], np.bincount(ends, minlength=len(grp))
            else:
                mon = mons[parent.name]
                deg = np.bincount(ends + getattr(grp, "start", 0), minlength=len(parent))
            i, t = np.asarray(mon.i[:]), np.asarray(mon.t_[:])
            sel = i[t >= t_from - 1e-12]
            total += float(deg[sel].sum())
            nspikes += int(len(sel))
    return total, nspikes


def _apply_prefs(b, workload, reset=False):
    defaults = {"devices.b200.construction": "reference"}
    for key, value in defaults.items():
        b.prefs[key] = value
    if not reset:
        for key, value in WORKLOAD_PREFS.get(workload, {}).items():
            b.prefs[key] = value


def run_b200(args, rank, world, workload=None, steps=None, warmup=None, sim_steps=None):
    """Build + run one workload on the b200 device; returns the measurements of this rank."""
    b = _import_brian()
    workload = workload or args.workload
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    if world > 1:
        _barrier(world)   # initialises the process group: the device shards over its ranks
    sim_steps = sim_steps or args.sim_steps or DEFAULT_SIM_STEPS.get(workload, 2000)
    n_runs = warmup + steps
    directory = _project_dir(f"bench_{workload}_w{1 if args.replicas else world}", rank)
    t_build0 = time.time()
    b.prefs["devices.b200.multi_gpu"] = not args.replicas
    b.prefs["devices.b200.persistent"] = not args.stepwise
    b.prefs["devices.b200.profile_phases"] = bool(args.phases)
    b.prefs["devices.b200.libm"] = args.libm
    if args.ctas_per_sm:
        b.prefs["devices.b200.ctas_per_sm"] = args.ctas_per_sm
    if args.grid:
        b.prefs["devices.b200.grid"] = args.grid
    _apply_prefs(b, workload)
    scale = 1 if args.replicas else world
    try:
        objs = _build_script(b, workload, "b200", directory, sim_steps, n_runs, scale=scale)
        b.device.build(directory=directory, compile=True, run=False, with_output=False)
    finally:
        sharded = b.prefs["devices.b200.construction"] == "sharded"
        _apply_prefs(b, workload, reset=True)
    build_seconds = time.time() - t_build0

    _barrier(world)
    sampler = ClockSampler(index=int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    t_run0 = time.time()
    b.device.run(directory=directory, with_output=False)
    run_wall = time.time() - t_run0
    cnt = b.device.counter
    runs = int(cnt("runs"))
    assert runs == n_runs, (runs, n_runs)
    timed = range(warmup, n_runs)
    clocks = sampler.stop([(cnt(f"run{r}.t0_unix"), cnt(f"run{r}.t0_unix") + cnt(f"run{r}.wall_seconds"))
                           for r in timed])
    dev_s = sum(cnt(f"run{r}.device_seconds") for r in timed)
    per_run_ms = [round(1e3 * cnt(f"run{r}.device_seconds"), 3) for r in range(n_runs)]
    # end to end = what a user's run() call costs: (re)build of the pathway CSRs if the host
    # arrays changed + upload + step loop + download
    e2e_s = sum(cnt(f"run{r}.prepare_seconds") + cnt(f"run{r}.upload_seconds") + cnt(f"run{r}.wall_seconds")
                + cnt(f"run{r}.download_seconds") for r in timed)
    e2e_parts = {k: 1e3 * sum(cnt(f"run{r}.{k}_seconds") for r in timed) / max(len(timed), 1)
                 for k in ("prepare", "upload", "wall", "download")}
    events = sum(cnt(f"run{r}.events") for r in timed)
    nsteps = sum(cnt(f"run{r}.steps") for r in timed)
    persistent = all(cnt(f"run{r}.persistent") == 1 for r in timed)
    if args.phases and world > 1:
        n = max(cnt("polls"), 1.0)
        sys.stderr.write(f"POLL rank {rank}: {int(n)} sampled waits, spin {cnt('poll_cycles') / n / 1.965e3:.3f} us, "
                         f"acquire fence {cnt('fence_cycles') / n / 1.965e3:.3f} us per wait\n")
    if args.phases:
        total_steps = cnt("steps")
        for name, cycs in b.device.phase_profile(all_ctas=True):
            cols = " ".join(f"{c / total_steps / 1.965e3:8.3f}" for c in cycs)
            sys.stderr.write(f"PHASE r{rank} {name:55s} {cols}  us/step (CTA 0, 1/4, 3/4, last)\n")
    model, kwds, bytes_neuron, bytes_event = WORKLOADS[workload]
    n_neurons = _n_neurons(objs)
    # synapses: of the whole network (sharded construction: the ranks' counts are added up by the caller)
    n_syn_local = sum(int(np_len(o)) for o in objs.values() if isinstance(o, b.Synapses)) if not sharded \
        else int(cnt("connect_synapses"))
    return dict(dev_s=dev_s, e2e_s=e2e_s, events=events, timesteps=nsteps, persistent=persistent,
                h2d=cnt("h2d_bytes") / n_runs, d2h=cnt("d2h_bytes") / n_runs, launches=cnt("launches"),
                clocks=clocks, n_neurons=n_neurons, n_syn=n_syn_local, sharded=sharded,
                build_seconds=build_seconds, run_wall=run_wall,
                connect_seconds=cnt("connect_seconds"), prepare_seconds=cnt("prepare_seconds"),
                sim_steps=sim_steps, bytes_neuron=bytes_neuron, bytes_event=bytes_event, e2e_parts=e2e_parts,
                objs=objs, workload=workload, steps=steps, warmup=warmup, grid=int(cnt("grid")),
                per_run_ms=per_run_ms)


def _project_dir(name, rank):
    """In-tree project directory of this rank.  The generated sources do not depend on the rank:
    a rank without a directory of its own starts from a copy of rank 0's (if that was built
    ahead, e.g. by tools/prebuild_bench.py), so that `make` finds nothing to do."""
    import shutil

    base = os.path.join(ROOT, "brian2_b200", "_prebuilt")
    mine, first = os.path.join(base, f"{name}_r{rank}"), os.path.join(base, f"{name}_r0")
    if rank > 0 and not os.path.isdir(mine) and os.path.exists(os.path.join(first, "libb200_project.so")):
        shutil.copytree(first, mine, ignore=shutil.ignore_patterns("results", "libb200_project_run*"))
    return mine


def np_len(obj):
    try:
        return len(obj)
    except Exception:
        return 0


def run_reference(args, sim_steps, warmup, steps, threads, strict=False, workload=None, scale=1):
    """The reference's own CPU implementation of the path (cpp_standalone [+ OpenMP])."""
    b = _import_brian()
    workload = workload or args.workload
    n_runs = warmup + steps
    directory = tempfile.mkdtemp(prefix="b200_bench_ref_")
    _apply_prefs(b, workload, reset=True)
    objs = _build_script(b, workload, "cpp_standalone", directory, sim_steps, n_runs,
                         openmp_threads=threads, strict=strict, scale=scale)
    b.device.build(directory=directory, compile=True, run=True, with_output=False)
    with open(os.path.join(b.device.results_dir, "stdout.txt")) as f:
        times = [float(line.split()[1]) for line in f if line.startswith("B200BENCH_RUN")]
    assert len(times) == n_runs, (times, n_runs)
    loop_s = sum(times[warmup:])
    n_syn = sum(len(o) for o in objs.values() if isinstance(o, b.Synapses))
    if "spikes" in objs:
        events, nspikes = _outdegree_events(b, objs, warmup * sim_steps * 1e-4)
    else:   # SynapsesOnly: every source spikes every step, so every synapse carries one event per step
        events, nspikes = float(n_syn) * steps * sim_steps, None
    n_neurons = _n_neurons(objs)
    return dict(loop_s=loop_s, events=events, timesteps=steps * sim_steps, n_neurons=n_neurons,
                spikes=nspikes, n_syn=n_syn, objs=objs)


# ---------------------------------------------------------------------------------------------
# parity checks carried by the bench line (outside every timed region)
# ---------------------------------------------------------------------------------------------
def parity_prefix_vs_reference(args, r, n_check=200):
    """1 GPU: the first `n_check` timesteps of the network that was just benchmarked, run again on
    the reference's cpp_standalone (serial, strict floating-point flags): the spike trains (i, t)
    must be identical."""
    import numpy as np

    b = _import_brian()
    mon = r["objs"].get("spikes")
    if mon is None:
        return {"checked": False, "why": "workload has no SpikeMonitor"}
    t_end = n_check * 1e-4 - 1e-9
    i_dev, t_dev = np.asarray(mon.i[:]), np.asarray(mon.t_[:])
    sel = t_dev < t_end
    i_dev, t_dev = i_dev[sel].copy(), t_dev[sel].copy()
    ref = run_reference(args, n_check, 0, 1, 0, strict=True, workload=r["workload"])
    mon_ref = ref["objs"]["spikes"]
    i_ref, t_ref = np.asarray(mon_ref.i[:]), np.asarray(mon_ref.t_[:])
    same = bool(i_dev.shape == i_ref.shape and np.array_equal(i_dev, i_ref) and np.array_equal(t_dev, t_ref))
    first_diff = None
    if not same:
        n = min(len(i_dev), len(i_ref))
        bad = np.nonzero((i_dev[:n] != i_ref[:n]) | (t_dev[:n] != t_ref[:n]))[0]
        first_diff = float(t_ref[bad[0]]) if len(bad) else float(min(t_dev[n - 1] if n else 0.0, t_ref[n - 1] if n else 0.0))
    out = {"checked": True, "identical": same, "timesteps": n_check, "spikes_compared": int(len(i_ref)),
           "against": "the reference's cpp_standalone (serial, -O3 -ffp-contract=off) on the same network, "
                      "same seed: spike trains (i, t) of the first timesteps", "first_difference_t": first_diff}
    try:
        # the reference's arrays must be read while its device is still the active one
        ref_state = {(g, v): np.array(getattr(ref["objs"][g], v + "_")[:], dtype=np.float64)
                     for g, v in ref["objs"]["state"]}
        out["state_check"] = state_bits_vs_reference(args, r["workload"], ref_state, (i_ref, t_ref), n_check)
    except Exception as ex:   # never hide the measurement
        out["state_check"] = {"checked": False, "why": f"{type(ex).__name__}: {ex}"}
    return out


def state_bits_vs_reference(args, workload, ref_state, ref_train, n_check):
    """The same `n_check` timesteps once more on the device with `prefs.devices.b200.libm = 'glibc'`
    (exp / expm1 / pow with the arithmetic of the host's glibc, csrc/b200_glibc_math.cuh): every
    state variable of every neuron must then have the bits of the reference's cpp_standalone run
    (`ref_state`, `ref_train`: final state and spike train of the strict, serial run of
    `parity_prefix_vs_reference`)."""
    import numpy as np

    b = _import_brian()
    directory = _project_dir(f"bench_state_{workload}", 0)
    b.prefs["devices.b200.multi_gpu"] = False
    b.prefs["devices.b200.libm"] = "glibc"
    _apply_prefs(b, workload)
    try:
        objs = _build_script(b, workload, "b200", directory, n_check, 1)
        b.device.build(directory=directory, compile=True, run=True, with_output=False)
    finally:
        b.prefs["devices.b200.libm"] = "cuda"
        b.prefs["devices.b200.multi_gpu"] = not args.replicas
        _apply_prefs(b, workload, reset=True)
    report = {"checked": True, "libm": "glibc", "timesteps": n_check, "variables": {}}
    ok = True
    for group, var in objs["state"]:
        dev = np.ascontiguousarray(getattr(objs[group], var + "_")[:], dtype=np.float64)
        ref = ref_state[(group, var)]
        differ = int((dev.view(np.uint64) != ref.view(np.uint64)).sum()) if dev.shape == ref.shape else dev.size
        report["variables"][f"{group}.{var}"] = {"values": int(dev.size), "not_bit_identical": differ}
        ok = ok and differ == 0
    mon = objs.get("spikes")
    if mon is not None:
        train_same = bool(np.array_equal(np.asarray(mon.i[:]), ref_train[0])
                          and np.array_equal(np.asarray(mon.t_[:]), ref_train[1]))
        report["spike_train_identical"] = train_same
        ok = ok and train_same
    report["bit_identical"] = bool(ok)
    report["against"] = "final state of every neuron after the same timesteps on cpp_standalone (and its spike train)"
    return report


def parity_multi_gpu(args, rank, world):
    """N > 1: (a) golden cases (reference connectivity) partitioned over the N ranks must reproduce
    the reference's spike trains; (b) a reduced Brunel network built per rank on the device must
    give, on N GPUs, exactly what rank 0 computes alone."""
    import numpy as np

    b = _import_brian()
    import models
    from golden.make_golden import CASES

    out = {"checked": True, "n_gpus": world, "cases": {}}
    ok = True
    for case in ("brunel_hetero", "cuba_1000"):
        model, kwds = CASES[case]
        d = _project_dir(f"bench_parity_{case}", rank)
        objs, res = models.run_model(b, model, "b200", d, **kwds)
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))
        same = all(np.array_equal(gold[k], res[k]) for k in gold.files)
        out["cases"][case] = {"identical_to_reference_golden": bool(same), "spikes": int(len(res["spikes_i"]))}
        ok = ok and same
        _barrier(world)
    # sharded construction: N ranks against one rank
    model, kwds = "brunel", dict(N_E=8000, epsilon=0.0125, duration=0.03)
    runs = {}
    for mode in ("ranks", "single"):
        if mode == "single" and rank != 0:
            continue
        d = _project_dir("bench_parity_sharded", rank)
        try:
            objs, res = models.run_model(b, model, "b200", d,
                                         prefs_update={"devices.b200.construction": "sharded",
                                                       "devices.b200.multi_gpu": mode == "ranks"}, **kwds)
        finally:
            b.prefs["devices.b200.construction"] = "reference"
            b.prefs["devices.b200.multi_gpu"] = not args.replicas
        res["exc_i"], res["exc_j"] = np.asarray(objs["exc"].i[:]), np.asarray(objs["exc"].j[:])
        res["exc_delay"] = np.asarray(objs["exc"].delay_[:])
        runs[mode] = res
    if rank == 0:
        keys = [k for k in runs["single"] if k != "last_run_time"]
        same = all(runs["single"][k].shape == runs["ranks"][k].shape and
                   np.array_equal(runs["single"][k], runs["ranks"][k]) for k in keys)
        out["cases"]["brunel_10k_sharded_construction"] = {
            "identical_to_1_gpu": bool(same), "spikes": int(len(runs["single"]["spikes_i"])),
            "synapses": int(runs["single"]["exc_nsyn"][0] + runs["single"]["inh_nsyn"][0])}
        ok = ok and same
    _barrier(world)
    out["identical"] = bool(ok)
    return out


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    return rank, world


_PG = {"init": False}


def _barrier(world):
    if world <= 1:
        return
    import torch
    import torch.distributed as dist

    if not _PG["init"]:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        # NCCL announces its version on stdout when the first communicator is created: keep
        # stdout for the one JSON line of the contract
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl")
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        _PG["init"] = True
    dist.barrier()
    torch.cuda.synchronize()


def _allreduce(values, op, world):
    if world <= 1:
        return values
    import torch
    import torch.distributed as dist

    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return t.tolist()


def _config(workload, n_neurons, n_syn, sharded):
    model, kwds, _, _ = WORKLOADS[workload]
    return {
        "workload": f"{workload}: {model} {kwds}, dt=0.1ms, fp64, "
                    + ("SpikeMonitor + StateMonitor of 3 voltage traces" if model == "cobahh" else
                       "no monitors" if model == "synapses_only" else "SpikeMonitor(s)"),
        "neurons": int(n_neurons), "synapses": int(n_syn),
        "connectivity": ("created per rank on the device (one Philox stream per source row, "
                         "csrc/b200_connect.cuh); statistically equivalent to the reference's connect()"
                         if sharded else
                         "reference Synapses.connect (host, mt19937), identical on both arms"),
        "l2": "no explicit flush: a bench step is one run() of thousands of simulation timesteps over the "
              "same state and CSR (a simulation re-reads its working set every timestep by nature; whether "
              "it is L2 resident is a property of the workload: COBAHH-256k state 14 MB + CSR 82 MB vs "
              "126 MB of L2); every timed step starts after the H2D upload of the state arrays",
    }


def _b200_line(args, r, world, hbm_peak, peak_src):
    """Whole-job JSON line of one measured workload (call on every rank; returns None off rank 0)."""
    rank, _ = _dist()
    dev_s, e2e_s = _allreduce([r["dev_s"], r["e2e_s"]], "MAX", world)
    sums = [r["events"]] + ([float(r["n_syn"])] if r["sharded"] else [])
    sums = _allreduce(sums, "SUM", world)
    events = sums[0]
    n_syn = int(sums[1]) if r["sharded"] else r["n_syn"]
    if rank != 0:
        return None
    steps, warmup = r["steps"], r["warmup"]
    timesteps = r["timesteps"]
    kw = WORKLOADS[r["workload"]][1]
    weak = "N" in kw or "N_E" in kw        # models without a size rule are sharded as they are
    bytes_neuron, bytes_event = r["bytes_neuron"], r["bytes_event"]
    # per-GPU roofline (rank 0): the neurons it owns and the synaptic events it delivered
    n_owned = r["n_neurons"] / (1 if (args.replicas or world == 1) else world)
    algo_bytes = timesteps * n_owned * bytes_neuron + r["events"] * bytes_event
    achieved = algo_bytes / r["dev_s"] / 1e9
    prop_bytes = r["events"] * bytes_event
    line = {
        "metric": "synaptic_events_per_s", "value": events / dev_s, "unit": "events/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dev_s / steps,
        "ms_per_run_rank0": r["per_run_ms"],
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": _config(r["workload"], r["n_neurons"], n_syn, r["sharded"]),
        "timesteps_per_step": r["sim_steps"],
        "parallelism": "1 GPU" if world == 1 else (
            f"{world} independent replicas (one per GPU)" if args.replicas else
            (f"one network of {world}x the neurons (same synapses per neuron)" if weak else
             "the same network (fixed size: strong scaling)") +
            f" partitioned by postsynaptic neuron over {world} GPUs; spike lists exchanged by NVLink peer "
            f"stores inside the persistent kernel"),
        "execution": ("persistent cooperative step kernel" if r["persistent"] else "one launch per code object")
                     + f", {r['grid']} CTAs x 512 threads"
                     + ("; exp/expm1/pow with the host glibc's arithmetic (prefs.devices.b200.libm = 'glibc')"
                        if args.libm == "glibc" else ""),
        "host_seconds": {"codegen_and_build": round(r["build_seconds"], 1), "device_run_call": round(r["run_wall"], 1),
                         "synapse_creation_on_device": round(r["connect_seconds"], 3),
                         "pathway_csr_build": round(r["prepare_seconds"], 2)},
        "realtime_factor": timesteps * 1e-4 / dev_s,
        "us_per_timestep": 1e6 * dev_s / max(timesteps, 1),
        "events_per_timestep": events / max(timesteps, 1),
        "clocks": r["clocks"],
        "e2e": {"value": events / e2e_s, "unit": "events/s", "h2d_bytes_per_step": r["h2d"],
                "d2h_bytes_per_step": r["d2h"],
                "ms_per_step": {"pathway_csr_build": r["e2e_parts"]["prepare"], "upload": r["e2e_parts"]["upload"],
                                "loop_host_clock": r["e2e_parts"]["wall"], "download": r["e2e_parts"]["download"]}},
        "gpu_launches": int(r["launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": _ncu_traffic(r["workload"], r["sim_steps"]),
                     "traffic_source": "profiles/ncu_traffic.json: DRAM bytes of an `ncu --set full` capture of "
                                       "the same workload, scaled to the timesteps of a launch (not this run)",
                     "algorithmic_bytes_per_launch": algo_bytes / max(steps, 1),
                     "peak_source": peak_src,
                     "kernel": "persistent step kernel (stateupdate+threshold+propagation+monitors)",
                     "limiter": LIMITERS.get(r["workload"].split("_")[0], "latency: dependent L2 round trips + grid barriers"),
                     "algorithmic_bytes": f"{bytes_neuron} B/neuron-step + {bytes_event} B/event",
                     "propagation_only": {"achieved": prop_bytes / r["dev_s"] / 1e9,
                                          "frac": prop_bytes / r["dev_s"] / 1e9 / hbm_peak,
                                          "note": "events x bytes/event over the WHOLE step time (state update, "
                                                  "barriers and monitors included), per GPU"}},
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--sim-steps", type=int, default=0, help="simulation timesteps per bench step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="default run: skip the extra configurations (Brunel) and the parity checks")
    ap.add_argument("--no-configs", action="store_true",
                    help="default run: keep the parity checks, skip the extra configurations (Brunel)")
    ap.add_argument("--replicas", action="store_true",
                    help="N>1: every rank simulates its own copy of the network instead of sharding one "
                         "N-times larger network over the ranks")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="tuning aid: prefs.devices.b200.ctas_per_sm")
    ap.add_argument("--grid", type=int, default=0, help="tuning aid: prefs.devices.b200.grid (max CTAs)")
    ap.add_argument("--libm", default="cuda", choices=["cuda", "glibc"],
                    help="prefs.devices.b200.libm: arithmetic of the device's exp/expm1/pow (glibc = bit-identical "
                         "state, slower state update)")
    ap.add_argument("--phases", action="store_true",
                    help="profiling aid: per-code-object cycle counters inside the persistent kernel")
    ap.add_argument("--stepwise", action="store_true",
                    help="profiling aid: one kernel launch per code object per step (not the product mode)")
    args = ap.parse_args()
    default_run = args.workload is None
    if default_run:
        args.workload = "cobahh_256k"
    rank, world = _dist()
    hbm_peak, peak_src = _peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        scale = 1 if args.replicas else max(1, args.gpus)
        sim_steps = args.sim_steps or max(10, REFERENCE_SIM_STEPS.get(args.workload, 500) // scale)
        r = run_reference(args, sim_steps, args.warmup, args.steps, threads, strict=False, scale=scale)
        value = r["events"] / r["loop_s"]
        line = {
            "impl": "reference", "metric": "synaptic_events_per_s", "value": value, "unit": "events/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["loop_s"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(args.workload, r["n_neurons"], r["n_syn"], False),
            "timesteps_per_step": sim_steps,
            "realtime_factor": r["timesteps"] * 1e-4 / r["loop_s"],
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": "reference",
                             "sample": f"{args.steps} x {sim_steps} timesteps of the same network "
                                       f"({r['n_neurons']} neurons, {r['n_syn']} synapses), "
                                       f"cpp_standalone + OpenMP({threads}), reference default flags"},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    r = run_b200(args, rank, world)
    line = _b200_line(args, r, world, hbm_peak, peak_src)
    extras, parity = [], None
    if default_run and not args.no_extra:
        # parity of the benchmarked configuration (1 GPU) / of the partitioned execution (N GPUs)
        try:
            parity = parity_prefix_vs_reference(args, r) if world == 1 else parity_multi_gpu(args, rank, world)
        except Exception as ex:   # a failed check must be visible, never hide the measurement
            parity = {"checked": False, "why": f"{type(ex).__name__}: {ex}"}
        for name in ([] if args.no_configs else EXTRA_CONFIGS):
            try:
                rx = run_b200(args, rank, world, workload=name, steps=min(args.steps, 10), warmup=min(args.warmup, 3))
                extra = _b200_line(args, rx, world, hbm_peak, peak_src)
            except Exception as ex:
                extra = {"config": {"workload": name}, "error": f"{type(ex).__name__}: {ex}"}
            if extra is not None:
                extras.append(extra)
    if rank != 0:
        return
    if parity is not None:
        line["parity_check"] = parity
    if extras:
        line["configs"] = extras
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref_steps = 50 if "256k" in args.workload else REFERENCE_SIM_STEPS.get(args.workload, 2000)
        try:
            ref = run_reference(args, ref_steps, 1, 2, threads, strict=False)
            line["cpu_baseline"] = {
                "value": ref["events"] / ref["loop_s"], "unit": "events/s", "cores": threads,
                "kind": "reference",
                "sample": f"2 x {ref_steps} timesteps (after 1 warm-up) of the same network on "
                          f"cpp_standalone + OpenMP({threads}), reference default flags",
                "realtime_factor": ref["timesteps"] * 1e-4 / ref["loop_s"],
            }
        except Exception as ex:  # the baseline must never hide the measurement
            line["cpu_baseline"] = {"value": None, "unit": "events/s", "cores": threads, "kind": "reference",
                                    "sample": f"failed: {ex}"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
