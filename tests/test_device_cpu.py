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


def test_barrier_plan_of_cuba(cuba_project):
    """stateupdate -> threshold share the owned partition (no barrier); one barrier before the
    consumers of the spike list (compaction, monitor, pathways; the resetter only touches the
    CTA's own segment); one at the end of the step."""
    src = open(os.path.join(cuba_project, "b200_kernels.cu")).read()
    m = re.search(r"schedule: (.*)\n", src)
    assert m
    sched = m.group(1).strip()
    assert sched == ("cuba_P_stateupdater_codeobject cuba_P_spike_thresholder_codeobject | "
                     "cuba_spikes_codeobject cuba_Ce_pre_codeobject cuba_Ci_pre_codeobject "
                     "cuba_P_spike_resetter_codeobject compact_array_cuba_P__spikespace"), sched
    assert "grid barriers per step: 2" in src


def test_synaptic_effect_uses_atomics_not_rmw(cuba_project):
    code = open(os.path.join(cuba_project, "code_objects", "cuba_Ce_pre_codeobject.cuh")).read()
    assert "b200::atomic_add(&_ptr_array_cuba_P_ge[_postsynaptic_idx]" in code
    # the reference's sequential read-modify-write must not survive in device code
    dev = code.split("__device__ __forceinline__ void _dev_")[1].split("__global__")[0]
    assert "ge[_postsynaptic_idx] = ge" not in dev.replace("_ptr_array_cuba_P_", "")


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
    (the reference *front-end* under oracle/_ref is located by _brian2_path.py only)."""
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
