"""CPU-only checks of the b200 device: code generation + nvcc cross-compilation for sm_100a, the
exported C ABI, the barrier plan of the persistent kernel, and the "no CPU fallback" contract.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def cuba_project(brian):
    import __graft_entry__ as ge

    directory, objs = ge.build_project("cuba_1000", directory=os.path.join(ge.PREBUILT, "cpu_cuba_1000"))
    return directory


def _declared_symbols():
    header = open(os.path.join(ROOT, "include", "brian2_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", header)))


def test_header_and_binding_agree():
    from brian2_b200.capi import ABI_SYMBOLS

    assert _declared_symbols() == sorted(ABI_SYMBOLS)


def test_library_builds_and_exports_every_declared_symbol(cuba_project):
    lib = os.path.join(cuba_project, "libb200_project.so")
    assert os.path.exists(lib)
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    exported = set(re.findall(r"\bT (b200_\w+)", out))
    missing = [s for s in _declared_symbols() if s not in exported]
    assert not missing, missing
    # loads with ctypes (no compute call)
    handle = ctypes.CDLL(lib, mode=ctypes.RTLD_LOCAL)
    handle.b200_get_counter.restype = ctypes.c_double
    handle.b200_get_counter.argtypes = [ctypes.c_char_p]
    assert handle.b200_get_counter(b"runs") == 0.0
    assert handle.b200_get_counter(b"no-such-counter") == -1.0


def test_device_code_is_sm100a_sass(cuba_project):
    lib = os.path.join(cuba_project, "libb200_project.so")
    out = subprocess.check_output(["cuobjdump", "-lelf", lib], text=True)
    assert "sm_100a" in out, out


def _schedules(src):
    """{variant tag: (phases, barriers per step, has end-of-step barrier)} of the first plan."""
    out = {}
    for tag, sched in re.findall(r"//   \[(d\d)\] schedule: (.*)\n", src):
        if tag in out:
            continue
        end = sched.rstrip().endswith("|")
        # (a model built twice in one process gets `_1`-suffixed code object names)
        phases = [[re.sub(r"_codeobject_\d+", "_codeobject", w) for w in p.split()]
                  for p in sched.rstrip().rstrip("|").split("|")]
        n = int(re.search(rf"//   \[{tag}\] grid barriers per step: (\d+)", src).group(1))
        out[tag] = (phases, n, end)
    return out


def test_barrier_plan_of_cuba(cuba_project):
    """Phases of a CUBA step.  stateupdate -> threshold -> reset share the owned partition (no
    barrier: the resetter only touches the CTA's own segment); one barrier before the consumers
    of the spike list (monitor, the two deliveries, compaction); one at the end of the step (the
    deliveries accumulate `ge`/`gi` with reductions, the next state update reads them).  The
    apply passes of the two "dual" pathways sit next to their deliveries: they only do
    something when the rows are dense (owner-computes over target tiles)."""
    plans = _schedules(open(os.path.join(cuba_project, "b200_kernels.cu")).read())
    assert list(plans) == ["d0"]          # delays make no difference to this schedule
    phases, n, end = plans["d0"]
    assert phases == [
        ["cuba_P_stateupdater_codeobject", "cuba_P_spike_thresholder_codeobject", "cuba_P_spike_resetter_codeobject"],
        ["cuba_spikes_codeobject", "cuba_Ce_pre_codeobject", "cuba_Ce_pre_codeobject@apply",
         "cuba_Ci_pre_codeobject", "cuba_Ci_pre_codeobject@apply", "compact_array_cuba_P__spikespace"]], phases
    assert (n, end) == (2, True)


def test_synaptic_effect_uses_reductions_not_rmw(cuba_project):
    """`ge_post += we`: the delivery issues reductions (the reference's sequential
    read-modify-write must not survive in the scattered code); the same statement survives, as a
    plain loop, only in the element-private apply pass used for dense rows."""
    code = open(os.path.join(cuba_project, "code_objects", "cuba_Ce_pre_codeobject.cuh")).read()
    deliver = code.split("__device__ __forceinline__ void _dev_cuba_Ce_pre_codeobject(")[1].split("__global__")[0]
    assert "b200::atomic_add(&_ptr_array_cuba_P_ge[_postsynaptic_idx]" in deliver
    assert "ge[_postsynaptic_idx] = ge" not in deliver.replace("_ptr_array_cuba_P_", "")
    assert "if (_pw.tileptr) return;" in deliver
    apply = code.split("void _dev_cuba_Ce_pre_codeobject_apply(")[1].split("__global__")[0]
    assert "extern __shared__ int _b200_tile[];" in apply
    assert "for (int _b200_k = 0; _b200_k < _b200_n; ++_b200_k)" in apply
    assert "ge += we;" in apply and "_b200_hits" not in apply


def test_counted_pathway_code_of_brunel(brian):
    """`v_post += J` with `v` declared `(unless refractory)` reads target-side state: the delivery
    counts events per target with integer reductions (no gather of `not_refractory` per event)
    and the owner applies `if(not_refractory) v += J` once per counted event."""
    import __graft_entry__ as ge

    directory, _ = ge.build_project("brunel_hetero", directory=os.path.join(ge.PREBUILT, "cpu_brunel_hetero"),
                                    compile=False)
    code = open(os.path.join(directory, "code_objects", "brunel_exc_pre_codeobject.cuh")).read()
    deliver = code.split("__device__ __forceinline__ void _dev_brunel_exc_pre_codeobject(")[1].split("__global__")[0]
    assert "atomicAdd(_b200_hits + _b200_tgt_idx, 1);" in deliver
    assert "not_refractory" not in deliver.split("// scalar code")[1]
    apply = code.split("void _dev_brunel_exc_pre_codeobject_apply(")[1].split("__global__")[0]
    assert "const int _b200_n = __ldcg(_b200_hits + _b200_tgt_idx);" in apply
    assert "if(not_refractory)" in apply and "v += J;" in apply


def test_no_cpu_fallback_without_gpu(cuba_project):
    """Running the project where no CUDA device exists must fail loudly, not fall back."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from brian2_b200.capi import B200Library

    results = os.path.join(cuba_project, "results")
    os.makedirs(results, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(cuba_project)
    try:
        lib = B200Library(os.path.join(cuba_project, "libb200_project.so"), fresh_copy=True)
        status = lib.run_main(["--results_dir", results + "/"])
    finally:
        os.chdir(cwd)
    assert status != 0
    assert "no CUDA device" in lib.last_error()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under brian2_b200/ may import, link or run it
    (the reference *front-end* lives under baseline/_ref, located by _brian2_path.py)."""
    pkg = os.path.join(ROOT, "brian2_b200")
    offenders = []
    for dirpath, dirnames, filenames in os.walk(pkg):
        if "_prebuilt" in dirpath or "__pycache__" in dirpath:
            continue
        for fn in filenames:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(dirpath, fn)).read()
            if "hotpath_oracle" in text or re.search(r"\bimport\s+oracle\b|from\s+oracle\b", text):
                offenders.append(fn)
    assert not offenders, offenders


def test_host_only_arrays_are_not_uploaded(cuba_project):
    """`_synaptic_pre/_post` and the per-synapse `delay` are read by the host-side CSR build only
    (the CUDA templates take both ends of a synapse from the CSR): no upload of them is generated,
    while every state array of the neurons still is."""
    src = open(os.path.join(cuba_project, "b200_objects.cpp")).read()
    uploads = re.findall(r"b200::upload_(?:vector|array|records)\(_A_host\.(\w+)", src)
    assert "_array_cuba_P_v" in uploads and "_array_cuba_P_ge" in uploads
    assert not [u for u in uploads if "_synaptic_p" in u or u.endswith("_delay")], uploads


def _emulate_propagation(bn, blr, rows_len, nwarps, mode):
    """Python restatement of the work distribution of templates/synapses.cu for one group of
    delay bins: bn[b] rows (spikes) in bin b, every row padded to blr[b] lines of 32 slots,
    rows_len[b][s] real lengths.  Returns a {(bin, row, slot): visits} counter."""
    from collections import Counter

    visits = Counter()
    nb = len(bn)
    bexcl = np.concatenate([[0], np.cumsum([bn[b] * blr[b] for b in range(nb)])])
    rexcl = np.concatenate([[0], np.cumsum(bn)])
    nlines, nrows = int(bexcl[-1]), int(rexcl[-1])

    def line_share(a, b):           # the `while (_a < _b)` loop: lines [a, b) of the sequence
        while a < b:
            L = int(np.searchsorted(bexcl, a, side="right") - 1)
            L = min(L, nb - 1)
            while bn[L] * blr[L] == 0:      # the ballot picks the last bin whose offset <= a
                L += 1
            loc = a - int(bexcl[L])
            s, l0 = divmod(loc, blr[L])
            nl = min(blr[L] - l0, b - a)
            a += nl
            n = rows_len[L][s]
            # the rows of a bin lie back to back in the CSR: the row starts at an arbitrary slot and
            # the lines are the aligned 32-slot lines it touches ((len + 62) >> 5 always suffice)
            rbeg = sum(rows_len[L][:s]) + 7 * L
            rend = rbeg + n
            abeg = (rbeg & ~31) + 32 * l0
            end = min(rend, abeg + 32 * nl)
            for lane in range(32):
                k0 = abeg + lane
                if k0 < rbeg:
                    k0 += 32        # first line of the row: lanes in front of its start
                for k in range(k0, end, 32):
                    visits[(L, s, k - rbeg)] += 1

    if mode == "row":
        assert nrows <= nwarps
        k = nwarps // nrows
        for w in range(nwarps):
            r, part = divmod(w, k)
            if r >= nrows:
                continue
            L = int(np.searchsorted(rexcl, r, side="right") - 1)
            s = r - int(rexcl[L])
            per = -(-blr[L] // k)
            lo = min(blr[L], part * per)
            hi = min(blr[L], lo + per)
            a = int(bexcl[L]) + s * blr[L] + lo
            line_share(a, a + hi - lo)
    elif mode == "line":
        for w in range(nwarps):
            line_share(w * nlines // nwarps, (w + 1) * nlines // nwarps)
    elif mode == "ticket":
        t = 0
        while 32 * t < nlines:      # whoever draws ticket t handles lines [32 t, 32 t + 32)
            line_share(32 * t, min(32 * t + 32, nlines))
            t += 1
    elif mode == "gather":
        for w in range(nwarps):
            ra, rb = w * nrows // nwarps, (w + 1) * nrows // nwarps
            for r0 in range(ra, rb, 32):
                batch = []
                for r in range(r0, min(r0 + 32, rb)):
                    L = int(np.searchsorted(rexcl, r, side="right") - 1)
                    batch.append((L, r - int(rexcl[L])))
                lens = [rows_len[L][s] for L, s in batch]
                lexcl = np.concatenate([[0], np.cumsum(lens)])
                for slot in range(int(lexcl[-1])):
                    j = int(np.searchsorted(lexcl, slot, side="right") - 1)
                    L, s = batch[j]
                    visits[(L, s, slot - int(lexcl[j]))] += 1
    return visits


@pytest.mark.parametrize("mode", ["row", "line", "ticket", "gather"])
def test_propagation_work_distribution_covers_every_synapse_once(mode):
    """Every (delay bin, spiking row, synapse slot) is delivered exactly once, for ragged rows,
    empty bins and empty rows, in each of the four ways templates/synapses.cu deals the work
    out (the arithmetic is restated in `_emulate_propagation`)."""
    rng = np.random.RandomState({"row": 1, "line": 2, "ticket": 3, "gather": 4}[mode])
    for trial in range(20):
        nb = int(rng.randint(1, 7))
        nwarps = int(rng.choice([16, 48, 160]))
        maxlen = [int(rng.randint(0, 130 if mode != "gather" else 66)) for _ in range(nb)]
        blr = [max(1, (m + 62) >> 5) for m in maxlen]
        if mode == "row":
            bn = [int(rng.randint(0, max(1, nwarps // nb))) for _ in range(nb)]
        elif mode == "gather":
            bn = [int(rng.randint(0, 3 * nwarps)) for _ in range(nb)]
        else:
            bn = [int(rng.randint(0, 40)) for _ in range(nb)]
        if sum(bn) == 0:
            bn[0] = 1
        rows_len = [[int(rng.randint(0, m + 1)) for _ in range(n)] for n, m in zip(bn, maxlen)]
        visits = _emulate_propagation(bn, blr, rows_len, nwarps, mode)
        expected = {(b, s, k) for b in range(nb) for s in range(bn[b]) for k in range(rows_len[b][s])}
        assert set(visits) == expected, (mode, trial)
        assert all(v == 1 for v in visits.values()), (mode, trial)


@pytest.fixture(scope="module")
def ragged_project(brian):
    import __graft_entry__ as ge

    directory, objs = ge.build_project("ragged", directory=os.path.join(ge.PREBUILT, "cpu_ragged"))
    return directory


def test_generated_propagation_code_of_ragged_case(ragged_project):
    """Generator decisions for the two pathways of the `ragged` model (cross-compiled for sm_100a by
    the fixture): `x_post += w` is a pure scatter -- unrolled delivery, `w` preloaded with the
    index stream, both ends of the synapse taken from the CSR; the on_post code writes synaptic
    and presynaptic variables -- no unrolling, no preload, plain store to `w`."""
    pre = open(os.path.join(ragged_project, "code_objects", "rg_S_pre_codeobject.cuh")).read()
    post = open(os.path.join(ragged_project, "code_objects", "rg_S_post_codeobject.cuh")).read()
    assert "double _b200_rd_w[4];" in pre and "const double w = _b200_rd_w[_u];" in pre
    assert "const int32_t _postsynaptic_idx = _b200_tgt_idx;" in pre
    assert "b200::atomic_add(&_ptr_array_rg_neurons_x[_postsynaptic_idx]" in pre
    assert "_synaptic_post[_idx]" not in pre.split("_dev_rg_S_pre_codeobject")[1]
    assert "_b200_rd_" not in post
    assert "const int32_t _presynaptic_idx = _b200_tgt_idx;" in post     # on_post: the other end is pre
    assert "b200::atomic_add(&_ptr_array_rg_neurons_y[_presynaptic_idx]" in post
    assert "_ptr_array_rg_S_w[_idx] = w;" in post
    # spikes -> on_pre -> on_post (reads `w`, which on_pre ... does not write here: same phase is
    # fine; on_post writes `w` that on_pre reads -> barrier): two barriers, none at the end
    plans = _schedules(open(os.path.join(ragged_project, "b200_kernels.cu")).read())
    phases, n, end = plans["d0"]
    assert "rg_S_pre_codeobject" in phases[1] and phases[2] == ["rg_S_post_codeobject"], phases
    assert n == 2 and not end


def _plans_of(case):
    import __graft_entry__ as ge

    directory, _ = ge.build_project(case, directory=os.path.join(ge.PREBUILT, "cpu_" + case), compile=False)
    return _schedules(open(os.path.join(directory, "b200_kernels.cu")).read())


def test_barrier_plan_of_brunel(brian):
    """`v += J` (exc) and `v += -g*J` (inh) only read target-side data (`not_refractory`): both
    are counted, so they may deliver side by side and the owner applies exc, then inh, then the
    reset -- the reference's order, bit-identical `v`, no floating-point atomics.  With the
    heterogeneous delays of the model (>= 1 step, 'd1') the deliveries overlap the state update
    and ONE grid barrier per step is left (round 1: four)."""
    plans = _plans_of("brunel_hetero")
    phases, n, end = plans["d0"]
    assert [len(p) for p in phases] == [2, 5, 3] and (n, end) == (2, False), phases
    assert phases[2] == ["brunel_exc_pre_codeobject@apply", "brunel_inh_pre_codeobject@apply",
                         "brunel_neurons_spike_resetter_codeobject"]
    phases, n, end = plans["d1"]
    assert phases[0] == ["brunel_neurons_stateupdater_codeobject", "brunel_neurons_spike_thresholder_codeobject",
                         "brunel_exc_pre_codeobject", "brunel_inh_pre_codeobject"], phases
    assert phases[1][1:4] == ["brunel_exc_pre_codeobject@apply", "brunel_inh_pre_codeobject@apply",
                              "brunel_neurons_spike_resetter_codeobject"], phases
    assert (len(phases), n, end) == (2, 1, False)


def test_barrier_plan_of_stdp(brian):
    """on_pre (order -1) and on_post (order +1) of one Synapses object touch the same synaptic
    variables (`w`, `Apre`, `Apost`, `lastupdate`): a barrier separates them; both thresholders
    share the first phase with both state updaters (element-private chains); the next step's
    state update reads `ge`, which on_pre accumulates with atomics: end-of-step barrier."""
    phases, n, end = _plans_of("stdp_1000")["d0"]
    assert len(phases) == 3, phases
    assert "stdp_S_pre_codeobject" in phases[1] and phases[2][0] == "stdp_S_post_codeobject"
    assert {"stdp_inputs_stateupdater_codeobject", "stdp_neurons_spike_thresholder_codeobject"} <= set(phases[0])
    assert (n, end) == (3, True)


def _build_and_dlopen(brian, name, make_network, prefs_update=None):
    """Generate + cross-compile a project and dlopen the library (RTLD_NOW: every symbol the
    generated main() calls must be defined -- no compute call is made)."""
    import models

    b = brian
    directory = os.path.join(ROOT, "brian2_b200", "_prebuilt", "cpu_" + name)
    b.device.reinit()
    b.device.activate()
    for key, value in (prefs_update or {}).items():
        b.prefs[key] = value
    try:
        b.set_device("b200", directory=directory, build_on_run=False)
        b.prefs.codegen.cpp.extra_compile_args_gcc = list(models.STRICT_GCC_FLAGS)
        b.defaultclock.dt = 0.1 * b.ms
        make_network(b)
        b.device.build(directory=directory, compile=True, run=False, with_output=False)
    finally:
        b.prefs["devices.b200.construction"] = "reference"
    ctypes.CDLL(os.path.join(directory, "libb200_project.so"), mode=os.RTLD_NOW | os.RTLD_LOCAL)
    return directory


def test_project_with_several_run_calls_links(brian):
    """Every run() call re-creates its code objects (`..._codeobject_1`): identical sources are
    compiled once and aliased, including the apply pass of counted pathways."""
    import models

    def net(b):
        objs = models.brunel(b, N_E=800, epsilon=0.1, duration=0.0)
        for _ in range(2):
            objs["net"].run(0.005 * b.second, namespace={})

    directory = _build_and_dlopen(brian, "two_runs", net)
    src = open(os.path.join(directory, "b200_kernels.cu")).read()
    assert re.search(r"void _run_brunel_exc_pre_codeobject_\d+_apply\(\) "
                     r"\{ _run_brunel_exc_pre_codeobject(_\d+)?_apply\(\); \}", src)
    assert "shares its kernels" in src


def test_sharded_connect_kernels_compile(brian):
    """prefs.devices.b200.construction = 'sharded': connect() generator expressions become CUDA
    kernels (p < 0.25 jump sampling, p >= 0.25, conditions on i/j with subgroup offsets, range
    generators, one-to-one), and expressions assigned to synaptic variables use the per-synapse
    host generator."""
    def net(b):
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
        b.Network(G, H, S1, S2, S3).run(1 * b.ms, namespace={})

    directory = _build_and_dlopen(brian, "sharded_connect", net,
                                  prefs_update={"devices.b200.construction": "sharded"})
    kern = open(os.path.join(directory, "code_objects", "sc_cond_synapses_create_generator_codeobject.cuh")).read()
    assert "b200::CandidateIter _it;" in kern and "_it.init_sample(" in kern
    assert "const b200::IdentityIndex _ptr_array_sc_G_i{ 0 };" in kern
    init = open(os.path.join(directory, "code_objects", "sc_dense_group_variable_set_conditional_codeobject.cpp")).read()
    assert "b200::SynapseRng _b200_synrng_rand(" in init and "brian::_random_generators" not in init
    # the reference's host connect is not generated for these objects
    assert not os.path.exists(os.path.join(directory, "code_objects", "sc_cond_synapses_create_generator_codeobject.cpp"))


def _run_host_cpp(tmp_path, source):
    exe = str(tmp_path / os.path.splitext(source)[0])
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "brian2_b200", "csrc"),
                           "-I", os.path.join(cuda_home, "include"),
                           os.path.join(ROOT, "tests", "cuda", source), "-o", exe,
                           "-L", os.path.join(cuda_home, "lib64"), "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_per_synapse_host_rng_is_a_function_of_the_synapse(tmp_path):
    """`b200::SynapseRng` (csrc/b200_synrng.h, sharded construction): `rand()`/`randn()` in an
    expression assigned to a synaptic variable give a synapse the same value on every rank layout
    (whole network in one array, or split by postsynaptic neuron), duplicates of a (pre, post)
    pair draw different numbers, and the draws have the right moments."""
    out = _run_host_cpp(tmp_path, "synrng_test.cpp")
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr


def test_forward_csr_layout_on_the_host(tmp_path):
    """`b200::Pathway::build_forward_csr` (csrc/b200_host.h): the (source, delay bin) layout of the
    forward delivery, checked on the CPU against a brute-force grouping of random synapses (with
    and without the partition filter of multi-GPU runs)."""
    exe = str(tmp_path / "forward_csr_test")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "brian2_b200", "csrc"),
                           "-I", os.path.join(cuda_home, "include"),
                           os.path.join(ROOT, "tests", "cuda", "forward_csr_test.cpp"), "-o", exe,
                           "-L", os.path.join(cuda_home, "lib64"), "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr


def test_rebuilt_library_in_the_same_directory_is_not_the_stale_one(brian):
    """glibc hands back the already loaded object when a path is dlopen()ed again, even if the
    file was rebuilt in between -- the second simulation would silently run the first model's code.
    `B200Library` therefore loads a private copy for every path it has loaded before."""
    import __graft_entry__ as ge
    from brian2_b200.capi import B200Library

    directory = os.path.join(ge.PREBUILT, "cpu_rebuild")
    ge.build_project("cuba_1000", directory=directory)
    first = B200Library(os.path.join(directory, "libb200_project.so"))
    assert first.lib.b200_get_array_size(b"cuba_P.v") == 1000 * 8
    ge.build_project("synapses_only", directory=directory)        # another model, same path
    second = B200Library(os.path.join(directory, "libb200_project.so"))
    assert second.path != first.path
    assert second.lib.b200_get_array_size(b"cuba_P.v") == -1
    assert second.lib.b200_get_array_size(b"so_targets.w") == 2000 * 8
    assert first.lib.b200_get_array_size(b"cuba_P.v") == 1000 * 8   # the first handle is untouched
