"""Locate the (unmodified) Brian2 front-end the ``b200`` device plugs into.

The device is a *plugin*: equations, ``NeuronGroup``/``Synapses``/monitors, ``Synapses.connect``
and the state updaters are Brian2's own (BASELINE.json north_star).  If ``brian2`` is not
already importable, fall back to the scripted install of the unmodified reference under
``baseline/_ref`` (see ``baseline/install_ref.py``; git-ignored, travels to the GPU box with the snapshot).
"""
import importlib.util
import os
import sys

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO_ROOT, "baseline", "_ref")


def _select_host_compiler():
    """Generated host code is compiled with the system g++ (shared libstdc++).  The image's
    default CXX (/opt/gcc/bin/g++) links libstdc++ statically -- which crashes in iostream locale
    code once the library is dlopen()ed into Python -- and cannot link -fopenmp."""
    if os.environ.get("B200_KEEP_CXX"):
        return
    # only replace what is unset or known to be unusable; any other choice of the user stands
    for var, good in (("CXX", "/usr/bin/g++"), ("CC", "/usr/bin/gcc")):
        current = os.environ.get(var, "")
        if os.path.exists(good) and (not current or current.startswith("/opt/gcc/")):
            os.environ[var] = good


def ensure_brian2_importable():
    _select_host_compiler()
    if "brian2" in sys.modules:
        return
    if importlib.util.find_spec("brian2") is not None:
        return
    if os.path.isdir(os.path.join(REF_DIR, "brian2")):
        sys.path.insert(0, REF_DIR)
        return
    raise ImportError(
        "brian2 is not importable and baseline/_ref is not populated; run "
        "`python baseline/install_ref.py` (needs /root/reference) or install brian2"
    )
