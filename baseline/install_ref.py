"""Recipe that installs the UNMODIFIED reference (brian-team/brian2) into ``baseline/_ref/``.

``baseline/_ref`` is git-ignored (never part of history) but is *not*
gpurun-ignored, so it travels to the GPU box exactly like our own built ``.so`` files.  It is
used for three things and nothing else:

* as the front-end the ``b200`` device plugs into (equations, ``Synapses.connect`` ... are the
  reference's own, untouched -- BASELINE.json north_star),
* as the parity oracle that generates the golden vectors (``set_device('cpp_standalone')``, serial,
  strict flags; tests/golden/make_golden.py), and
* as the CPU baseline arm of ``bench.py`` (``cpp_standalone`` + OpenMP on the box's host cores).

Recipe (verified): copy the package where it lies under /root/reference,
build its two Cython extensions in place (``setup.py:37-52``: cythonspikequeue,
cythondynamicarray), and expose pip's vendored pyparsing 3.1 (the image has no ``pyparsing``;
``pyproject.toml:14`` needs ``>=3``).  No reference source is copied into tracked files.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("BRIAN2_REFERENCE", "/root/reference")


def installed() -> bool:
    import glob

    return bool(glob.glob(os.path.join(DST, "brian2", "synapses", "cythonspikequeue*.so"))) and \
        os.path.isdir(os.path.join(DST, "pyparsing"))


def install(force: bool = False) -> str:
    if installed() and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "brian2")):
        raise RuntimeError(f"reference not found at {SRC} and {DST} is not populated")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for item in ("brian2", "setup.py", "pyproject.toml", "README.md"):
        s = os.path.join(SRC, item)
        d = os.path.join(DST, item)
        if os.path.isdir(s):
            shutil.copytree(s, d)
        else:
            shutil.copy(s, d)
    subprocess.check_call(["chmod", "-R", "u+w", DST])
    env = dict(os.environ, CXX="/usr/bin/g++", CC="/usr/bin/gcc")
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=DST, env=env,
                          stdout=subprocess.DEVNULL)
    import pip

    vend = os.path.join(os.path.dirname(pip.__file__), "_vendor", "pyparsing")
    # copy (not symlink) so it survives the trip to the GPU box unchanged
    shutil.copytree(vend, os.path.join(DST, "pyparsing"))
    shutil.rmtree(os.path.join(DST, "build"), ignore_errors=True)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
