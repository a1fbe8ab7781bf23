"""`prefs.devices.b200.libm = 'glibc'` (csrc/b200_glibc_math.cuh, brian2_b200/libm_tables.py):
the device's exp / expm1 / log / pow (and tanh / sinh / cosh / sin / cos) must be the host glibc's functions bit for bit, because the oracle
of this path -- the reference's cpp_standalone build -- calls exactly those.  Everything that can
be checked without a GPU is checked here: the tables found in the host's libm mean what the
algorithm assumes, the restated arithmetic (compiled for the host from the same header the device
compiles) returns glibc's bits over > 10^7 arguments per function, and a Hodgkin-Huxley project
cross-compiles with it.  The device run itself: tests/test_parity_gpu.py
(`test_cobahh_state_bit_exact_with_glibc_math`)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tables():
    from brian2_b200 import libm_tables

    if not libm_tables.host_has_fma_variants():
        pytest.skip("host without FMA/AVX2: its glibc runs other variants of exp/pow/expm1")
    return libm_tables.read_tables()


def test_tables_of_the_host_libm_mean_what_the_algorithms_assume(tables):
    """exp: entry i holds H = RN(2^(i/128)) (minus i << 45 in the exponent field) and T with
    H (1 + T) = 2^(i/128) to ~2^-107; pow: (1/c, log c high, log c low) with
    log(c) = -log(1/c) to ~2^-98 and 1/c a short number (so that z/c - 1 is exact in one fma)."""
    mp = pytest.importorskip("mpmath")
    from brian2_b200.libm_tables import _dbl

    mp.mp.prec = 240
    assert len(tables["exp_tab"]) == 256 and len(tables["pow_tab"]) == 384 and len(tables["log_tab"]) == 256
    for k in range(110):                 # sin/cos: value of sin(k/128), cos(k/128) as high + low
        sn, ssn, cs, ccs = (_dbl(v) for v in tables["sincos_tab"][4 * k:4 * k + 4])
        xk = mp.mpf(k) / 128
        assert abs(mp.mpf(sn) + mp.mpf(ssn) - mp.sin(xk)) < mp.mpf(2) ** -100, k
        assert abs(mp.mpf(cs) + mp.mpf(ccs) - mp.cos(xk)) < mp.mpf(2) ** -100, k
    for i in range(128):
        H = _dbl(tables["exp_tab"][2 * i + 1] + (i << 45))
        T = _dbl(tables["exp_tab"][2 * i])
        exact = mp.mpf(2) ** (mp.mpf(i) / 128)
        assert H == float(exact), i
        assert abs(mp.mpf(H) * (1 + mp.mpf(T)) - exact) / exact < mp.mpf(2) ** -100, i
    for i in range(128):
        invc, logc, tail = (_dbl(v) for v in tables["pow_tab"][3 * i:3 * i + 3])
        assert abs(mp.mpf(logc) + mp.mpf(tail) + mp.log(mp.mpf(invc))) < mp.mpf(2) ** -90, i
        assert tables["pow_tab"][3 * i] & ((1 << 40) - 1) == 0, (i, "1/c must have a short mantissa")
    for i in range(128):                 # log: (1/c, log c), log c rounded to 2^-43 and c chosen close to it
        invc, logc = (_dbl(v) for v in tables["log_tab"][2 * i:2 * i + 2])
        assert abs(mp.mpf(logc) + mp.log(mp.mpf(invc))) < mp.mpf(2) ** -60, i
        assert mp.mpf(logc) * 2 ** 43 == mp.floor(mp.mpf(logc) * 2 ** 43), i
    # subintervals [c_i (1 - 1/256), c_i (1 + 1/256)) tile [0x1.69555p-1, 0x1.69555p0)
    centres = sorted(1.0 / _dbl(tables["pow_tab"][3 * i]) for i in range(128))
    assert 0.70 < centres[0] < 0.71 and 1.40 < centres[-1] < 1.42


def test_header_is_self_describing(tables, tmp_path):
    from brian2_b200 import libm_tables

    path = libm_tables.write_header(str(tmp_path), tables)
    text = open(path).read()
    assert tables["path"] in text and "#define B200_LIBM_EXP_TAB" in text and "#define B200_LIBM_POW_TAB" in text
    assert len(re.findall(r"0x[0-9a-f]{16}ull", text)) == 256 + 384 + 256 + 440
    stamp = os.path.getmtime(path)
    libm_tables.write_header(str(tmp_path), tables)          # unchanged content: not rewritten (make)
    assert os.path.getmtime(path) == stamp


def test_restated_functions_return_the_bits_of_the_host_glibc(tables, tmp_path):
    """tests/cuda/glibc_math_test.cpp: 57 argument distributions (Hodgkin-Huxley ranges, whole
    range, over/underflow, subnormal results, random bit patterns, special values), 4 * 10^6
    arguments each (> 10^7 per function), every result compared with the libm call the
    reference's C++ code makes.  NaNs compare equal to NaNs."""
    from brian2_b200 import libm_tables

    libm_tables.write_header(str(tmp_path), tables)
    exe = str(tmp_path / "glibc_math_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", str(tmp_path),
                           "-I", os.path.join(ROOT, "brian2_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda", "glibc_math_test.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "4000000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if "arguments" in l]
    assert len(lines) == 57 and all(l.rstrip().endswith(", 0 differ") for l in lines), out.stdout
    for fn in ("exp ", "expm1", "log ", "pow ", "tanh", "sinh", "cosh", "sin ", "cos "):
        assert sum(int(l.split()[-4]) for l in lines if l.startswith(fn)) > 10 ** 7, fn


def test_the_check_above_can_fail(tables, tmp_path):
    """Negative control: a polynomial coefficient that is off in its 29th bit is detected."""
    from brian2_b200 import libm_tables

    text = libm_tables.header_text(tables)
    c2 = float(tables["exp_k"][4]).hex()
    assert c2 in text
    broken = text.replace(c2, (float(tables["exp_k"][4]) * (1 + 2.0 ** -29)).hex())
    (tmp_path / "b200_libm_tables.h").write_text(broken)
    exe = str(tmp_path / "glibc_math_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", str(tmp_path),
                           "-I", os.path.join(ROOT, "brian2_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda", "glibc_math_test.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "100000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 1 and out.stdout.strip().endswith("FAIL")


def test_hodgkin_huxley_project_builds_with_glibc_math(brian):
    """The generated COBAHH project compiles for sm_100a with -DB200_GLIBC_MATH: the tables header
    is written into the project, exp/expm1/pow of the state update are the restated functions
    (table loads in the kernel, no call into CUDA's libdevice exp/pow slow paths needed for them),
    and the default build of the same model does not contain any of it."""
    import models
    import __graft_entry__ as ge

    directory = os.path.join(ge.PREBUILT, "cpu_cobahh_glibc")
    b = brian
    b.device.reinit()
    b.device.activate()
    b.prefs["devices.b200.libm"] = "glibc"
    try:
        b.set_device("b200", directory=directory, build_on_run=False)
        b.prefs.codegen.cpp.extra_compile_args_gcc = list(models.STRICT_GCC_FLAGS)
        b.defaultclock.dt = 0.1 * b.ms
        objs = models.MODELS["cobahh"](b, N=1000, duration=0.05)
        objs["net"].run(objs["duration"] * b.second, namespace={})
        b.device.build(directory=directory, compile=True, run=False, with_output=False)
    finally:
        b.prefs["devices.b200.libm"] = "cuda"
    assert os.path.exists(os.path.join(directory, "b200_libm_tables.h"))
    assert "-DB200_GLIBC_MATH" in open(os.path.join(directory, "makefile")).read()
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(directory, "libb200_project.so")],
                          capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    elf = subprocess.run(["cuobjdump", "-elf", os.path.join(directory, "libb200_project.so")],
                         capture_output=True, text=True).stdout
    assert "kExpTab" in elf and "kPowLogTab" in elf and "kLogTab" in elf
    with pytest.raises(Exception):
        b.prefs["devices.b200.libm"] = "fdlibm"
    assert b.prefs["devices.b200.libm"] == "cuda"


def test_generated_code_routes_every_libm_call_through_the_switchable_wrappers(brian):
    """Code generation only (tests/models.py: mathfuncs): every transcendental function of the
    in-loop code is rendered as its `_b200_*` wrapper (csrc/b200_functions.cuh), which is CUDA's
    function by default and glibc's arithmetic under -DB200_GLIBC_MATH; powers go through
    `_brian_pow`, `exp(a)**c` through `_b200_exp_pow`."""
    import __graft_entry__ as ge

    directory, _ = ge.build_project("mathfuncs", directory=os.path.join(ge.PREBUILT, "plan_mathfuncs"),
                                    compile=False, prefs_update={"devices.b200.libm": "glibc"})
    code = open(os.path.join(directory, "code_objects", "mf_rr_codeobject.cuh")).read()
    for call in ("_b200_exp(x)", "_b200_expm1(0.1 * x)", "_exprel(0.05 * x)", "_b200_log(p)", "_brian_pow(p, q)",
                 "_brian_pow(x, 2)", "_brian_pow(p, - 1)", "_brian_pow(w, 3)", "_b200_exp_pow(0.01 * x, 0.3)",
                 "_b200_tanh(0.1 * x)", "_b200_sinh(0.2 * x)", "_b200_cosh(0.2 * x)", "_b200_sin(x * p)",
                 "_b200_cos((x * p) * p)"):
        assert call in code, call
    assert not re.search(r"[^_a-z0-9](exp|expm1|log|sin|cos|tanh|sinh|cosh|pow)\(", code)
    functions = open(os.path.join(ROOT, "brian2_b200", "csrc", "b200_functions.cuh")).read()
    for fn in ("exp", "expm1", "log", "pow", "tanh", "sinh", "cosh", "sin", "cos"):
        assert f"b200g::{fn}(" in functions or f"B200_LIBM_WRAPPER({fn})" in functions, fn
