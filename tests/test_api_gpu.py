"""The boundary itself on a GPU: the C ABI of include/brian2_b200.h called through ctypes
(`b200_get_array/_set_array/_profiling/_request_stop`), and the reference's run-argument workflow
(`device.run(run_args=...)` re-runs a built project with other parameter values without
recompiling; reference tests: brian2/tests/test_cpp_standalone.py:712-992)."""
import os
import threading
import time

import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu


def _fresh(b, project_dir, **kwds):
    b.device.reinit()
    b.device.activate()
    b.set_device("b200", directory=project_dir, with_output=False, **kwds)
    b.prefs.codegen.cpp.extra_compile_args_gcc = list(models.STRICT_GCC_FLAGS)
    b.defaultclock.dt = 0.1 * b.ms


def test_run_args_change_parameters_without_recompile(brian, project_dir):
    """test_cpp_standalone.py:712-806 on the b200 device (fp64): scalar values, values from files
    and TimedArray contents are replaced at run time through `./main name=value` arguments --
    here the argv of b200_run_main."""
    b = brian
    _fresh(b, project_dir)
    nA, volt = b.nA, b.volt
    on_off = b.TimedArray([True, False, True], dt=b.defaultclock.dt, name="ra_on_off")
    stim = b.TimedArray(np.arange(30).reshape(3, 10) * nA, dt=b.defaultclock.dt, name="ra_stim")
    G = b.NeuronGroup(10, """x : 1 (constant)
                             v : volt (constant)
                             n : integer (constant)
                             b : boolean (constant)
                             s = int(ra_on_off(t))*ra_stim(t, i) : amp""", name="ra_neurons",
                      namespace=dict(ra_on_off=on_off, ra_stim=stim))
    G.x = np.arange(10)
    G.n = np.arange(10)
    G.b = np.arange(10) % 2 == 0
    G.v = np.arange(10) * volt
    mon = b.StateMonitor(G, "s", record=True, name="ra_mon")
    net = b.Network(G, mon)
    net.run(3 * b.defaultclock.dt, namespace={})
    mtime = os.path.getmtime(os.path.join(project_dir, "libb200_project.so"))
    assert np.array_equal(G.x[:], np.arange(10)) and np.array_equal(G.n[:], np.arange(10))
    rows = np.arange(30).reshape(3, 10).astype(float)
    np.testing.assert_allclose(np.asarray(mon.s_).T / 1e-9, rows * [[1], [0], [1]])

    b.device.run(run_args=["ra_neurons.x=5", "ra_neurons.v=3", "ra_neurons.n=17", "ra_neurons.b=True",
                           "ra_on_off.values=True"])
    assert np.array_equal(G.x[:], np.ones(10) * 5) and np.array_equal(G.n[:], np.ones(10) * 17)
    assert np.array_equal(G.b[:], np.ones(10, dtype=bool)) and np.array_equal(G.v_[:], np.ones(10) * 3)
    np.testing.assert_allclose(np.asarray(mon.s_).T / 1e-9, rows)

    ar = np.arange(10) * 2.0
    ar.astype(G.x.dtype).tofile(os.path.join(project_dir, "init_values_x1.dat"))
    ar.astype(G.n.dtype).tofile(os.path.join(project_dir, "init_values_n1.dat"))
    (2 * np.arange(30).reshape(3, 10) * 1e-9).astype(np.float64).tofile(os.path.join(project_dir, "init_stim.dat"))
    # dictionary syntax (test_cpp_standalone.py:809-893): variables as keys, arrays as values
    b.device.run(run_args={G.x: ar, G.n: ar.astype(G.n.dtype), stim: 2 * np.arange(30).reshape(3, 10) * nA,
                           on_off: True})
    assert np.array_equal(G.x[:], ar) and np.array_equal(G.n[:], ar)
    np.testing.assert_allclose(np.asarray(mon.s_).T / 1e-9, 2 * rows)
    assert os.path.getmtime(os.path.join(project_dir, "libb200_project.so")) == mtime    # no rebuild


def test_c_abi_arrays_profiling_and_counters(brian, project_dir):
    """b200_get_array / b200_set_array / b200_get_array_size / b200_profiling / b200_get_counter
    called directly on the loaded library after a profiled run."""
    b = brian
    _fresh(b, project_dir, build_on_run=False)
    objs = models.cuba(b, N=1000, p=0.08, duration=0.02)
    objs["net"].run(0.02 * b.second, namespace={}, profile=True)
    b.device.build(directory=project_dir, compile=True, run=True, with_output=False)
    lib = b.device._b200_library
    P = objs["P"]
    v = lib.get_array("cuba_P.v", np.float64)                 # by "<owner>.<variable>" ...
    assert np.array_equal(v, np.asarray(P.v_[:]))
    assert np.array_equal(lib.get_array("_array_cuba_P_v", np.float64), v)    # ... or by array name
    i = lib.get_array("cuba_spikes.i", np.int32)
    assert np.array_equal(i, np.asarray(objs["spikes"].i[:]))
    lib.set_array("cuba_P.ge", np.full(1000, 0.25))
    assert np.array_equal(lib.get_array("cuba_P.ge", np.float64), np.full(1000, 0.25))
    with pytest.raises(KeyError):
        lib.get_array("no_such.array", np.float64)
    prof = dict(lib.profiling())
    assert any("stateupdater" in name for name in prof) and all(sec >= 0.0 for sec in prof.values())
    assert sum(prof.values()) > 0.0
    info = dict(objs["net"].get_profiling_info())             # results/profiling_info.txt
    assert set(info) and all(float(t) >= 0 for t in info.values())
    assert lib.get_counter("steps") == 200 and lib.get_counter("launches") > 200     # stepwise: profiled
    assert lib.last_run_time() > 0 and lib.last_run_completed_fraction() == 1.0


def test_request_stop_ends_the_persistent_kernel_early(brian, project_dir):
    """b200_request_stop (the reference stops on SIGINT, main.cpp:38-51 -> Network::_globally_stopped):
    a host-mapped flag polled by the persistent kernel every 64 steps."""
    b = brian
    _fresh(b, project_dir, build_on_run=False)
    objs = models.cuba(b, N=1000, p=0.08, duration=0.0, monitor=False)
    objs["net"].run(100 * b.second, namespace={})              # 10^6 steps: seconds of GPU time
    b.device.build(directory=project_dir, compile=True, run=False, with_output=False)
    dev = b.get_device()            # (the real device object, not the `brian2.device` proxy)
    dev._b200_library = None

    def stopper():
        t0 = time.time()
        while dev._b200_library is None and time.time() - t0 < 60:
            time.sleep(0.01)
        lib = dev._b200_library
        while lib.get_counter("steps") <= 0 and time.time() - t0 < 120:    # the step loop is running
            time.sleep(0.01)
        lib.request_stop()

    th = threading.Thread(target=stopper)
    th.start()
    dev.run(directory=project_dir, with_output=False)
    th.join()
    assert 0.0 < dev._last_run_completed_fraction < 1.0
    assert 0 < dev.counter("steps") < 1e6
