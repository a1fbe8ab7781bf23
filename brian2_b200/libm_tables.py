"""
Constants and lookup tables of the host's glibc ``exp`` / ``pow``, read from the libm binary the
reference's ``cpp_standalone`` build links against.

``prefs.devices.b200.libm = 'glibc'`` makes the device evaluate ``exp``, ``expm1``, ``log`` and ``pow`` with
the arithmetic of the host's libm (csrc/b200_glibc_math.cuh), so that state variables are
bit-identical to a ``cpp_standalone`` run -- the oracle of this path.
glibc >= 2.28 computes exp/pow/log from small tables (``__exp_data``: 2^(i/128) as value + error
term; ``__pow_log_data``: 1/c, log(c) in two pieces for 128 subintervals of [0.71, 1.42);
``__log_data``: 1/c, log(c) for 128 subintervals of [0.69, 1.38); sin/cos use ``__sincostab``:
sin and cos of k/128 in two pieces) whose
entries were chosen by a search, not by a closed formula, so they cannot be recomputed here: they
are looked up in the ``.rodata`` of the very library the oracle calls, by content (each table
follows a run of constants with known values), checked structurally, and written into the
project directory as ``b200_libm_tables.h``.  Nothing of glibc is stored in this repository.

The x86-64 build of glibc selects its functions at load time; with FMA and AVX2 (every host this
package targets) the ``__exp_fma/__pow_fma/__log_fma/__expm1_fma`` variants run, whose contraction of
``a*b+c`` into fused operations is restated operation by operation in b200_glibc_math.cuh.
``tests/cuda/glibc_math_test.cpp`` compiles that header for the host and compares it bit by bit
with the real functions (tests/test_glibc_math_cpu.py).
"""
import ctypes
import math
import os
import struct

__all__ = ["find_libm", "read_tables", "header_text", "write_header", "host_has_fma_variants"]

_EXP_TABLE_BITS = 7
_N = 1 << _EXP_TABLE_BITS


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def _dbl(u):
    return struct.unpack("<d", struct.pack("<Q", u & 0xFFFFFFFFFFFFFFFF))[0]


def find_libm():
    """Path of the libm shared object loaded into this process (= the one a g++ build on this
    host links against)."""
    ctypes.CDLL("libm.so.6")
    with open("/proc/self/maps") as f:
        for line in f:
            path = line.rsplit(" ", 1)[-1].strip()
            base = os.path.basename(path)
            if base.startswith("libm.so") or (base.startswith("libm-") and base.endswith(".so")):
                return path
    raise RuntimeError("libm.so.6 is not mapped into this process")


def host_has_fma_variants():
    """True if glibc's load-time selection picks the FMA variants on this host (FMA + AVX2)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return "fma" in flags and "avx2" in flags
    except OSError:
        pass
    return False


def _rodata(path):
    """(bytes, file offset) of the .rodata section of an ELF64 little-endian file."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:6] != b"\x7fELF\x02\x01":
        raise RuntimeError(f"{path}: not a 64-bit little-endian ELF file")
    shoff = struct.unpack_from("<Q", blob, 0x28)[0]
    shentsize, shnum, shstrndx = struct.unpack_from("<HHH", blob, 0x3A)
    sections = []
    for i in range(shnum):
        name, _type, _flags, _addr, off, size = struct.unpack_from("<IIQQQQ", blob, shoff + i * shentsize)
        sections.append((name, off, size))
    _, stroff, strsize = sections[shstrndx]
    strtab = blob[stroff:stroff + strsize]
    for name, off, size in sections:
        end = strtab.index(b"\0", name)
        if strtab[name:end] == b".rodata":
            return blob[off:off + size], off
    raise RuntimeError(f"{path}: no .rodata section")


def _find_all(data, values):
    """Offsets of all 8-byte aligned occurrences of the doubles `values` in `data`."""
    pattern = b"".join(struct.pack("<d", v) for v in values)
    hits = []
    pos = data.find(pattern)
    while pos >= 0:
        if pos % 8 == 0:
            hits.append(pos)
        pos = data.find(pattern, pos + 1)
    return hits


def _find_run(data, values, what):
    """Offset of the unique 8-byte aligned occurrence of the doubles `values` in `data`."""
    hits = _find_all(data, values)
    if len(hits) != 1:
        raise RuntimeError(f"libm: {what} found {len(hits)} times (expected once); unsupported glibc build")
    return hits[0]


def read_tables(path=None):
    """Locate ``__exp_data`` and ``__pow_log_data`` in the host's libm.

    Returns a dict: ``exp_k`` (InvLn2N, Shift, NegLn2hiN, NegLn2loN, C2..C5 as doubles), ``exp_tab``
    (256 uint64: error term, scaled value per entry), ``pow_k`` (Ln2hi, Ln2lo, A0..A6), ``pow_tab``
    (128 x (invc, logc, logctail) as uint64), ``log_k`` (Ln2hi, Ln2lo, A0..A4, B0..B10), ``log_tab``
    (128 x (invc, logc) as uint64), ``sincos_tab`` (110 x (sin high, low, cos high, low) as uint64),
    ``path``.
    """
    path = path or find_libm()
    data, _ = _rodata(path)
    u64 = lambda off, n: list(struct.unpack_from(f"<{n}Q", data, off))
    f64 = lambda off, n: list(struct.unpack_from(f"<{n}d", data, off))

    # ---- __exp_data: {invln2N, shift, negln2hiN, negln2loN, poly[4], ..., tab[2*N]} ------------
    invln2n = float.fromhex("0x1.71547652b82fep0") * _N
    shift = float.fromhex("0x1.8p52")
    base = _find_run(data, [invln2n, shift], "exp constants")
    exp_k = f64(base, 8)
    if not (abs(exp_k[2] + math.log(2) / _N) < 1e-12 and abs(exp_k[3]) < 1e-12
            and abs(exp_k[4] - 0.5) < 1e-9 and abs(exp_k[5] - 1 / 6) < 1e-9
            and abs(exp_k[6] - 1 / 24) < 1e-6 and abs(exp_k[7] - 1 / 120) < 1e-6):
        raise RuntimeError("libm: unexpected layout of the exp constants")
    tab_off = None
    one = _bits(1.0)
    for off in range(base + 64, base + 64 + 1024, 8):
        a, b, _c, d = u64(off, 4)
        if a == 0 and b == one and abs(_dbl(d + (1 << (52 - _EXP_TABLE_BITS))) - 2.0 ** (1 / _N)) < 1e-15:
            tab_off = off
            break
    if tab_off is None:
        raise RuntimeError("libm: exp table not found")
    exp_tab = u64(tab_off, 2 * _N)
    for i in range(_N):
        value = _dbl(exp_tab[2 * i + 1] + (i << (52 - _EXP_TABLE_BITS)))
        tail = _dbl(exp_tab[2 * i])
        exact = 2.0 ** (i / _N)
        if not (abs(value - exact) <= 4e-16 * exact and abs(tail) < 2.0 ** -53):
            raise RuntimeError(f"libm: exp table entry {i} fails its check")

    # ---- __pow_log_data: {ln2hi, ln2lo, poly[7], tab[N]{invc, pad, logc, logctail}} -----------
    ln2hi = float.fromhex("0x1.62e42fefa3800p-1")
    ln2lo = float.fromhex("0x1.ef35793c76730p-45")
    pbase = _find_run(data, [ln2hi, ln2lo, -0.5], "pow-log constants")
    pow_k = f64(pbase, 9)
    if not (abs(pow_k[3] + 2 / 3) < 1e-9 and abs(pow_k[4] - 0.5) < 1e-9):   # scaled by -2, 4/-2, ...
        raise RuntimeError("libm: unexpected layout of the pow-log constants")
    ptab = pbase + 9 * 8
    pow_tab = []
    for i in range(_N):
        invc, _pad, logc, logctail = f64(ptab + 32 * i, 4)
        if not (0.70 < invc < 1.42 and abs(logc + logctail + math.log(invc)) < 1e-13
                and abs(logctail) < 2.0 ** -43):
            raise RuntimeError(f"libm: pow-log table entry {i} fails its check")
        raw = u64(ptab + 32 * i, 4)
        pow_tab += [raw[0], raw[2], raw[3]]

    # ---- __log_data: {ln2hi, ln2lo, poly[5], poly1[11], tab[N]{invc, logc}, ...} --------------
    # same ln2 split as pow; told apart by poly[0] = -0.5 - 1 ulp and poly1[0] = -0.5 exactly
    a0 = float.fromhex("-0x1.0000000000001p-1")
    lbase = [h for h in _find_all(data, [ln2hi, ln2lo, a0]) if f64(h + 7 * 8, 1)[0] == -0.5]
    if len(lbase) != 1:
        raise RuntimeError(f"libm: log constants found {len(lbase)} times (expected once)")
    lbase = lbase[0]
    log_k = f64(lbase, 18)            # ln2hi, ln2lo, A0..A4, B0..B10
    if not (abs(log_k[3] - 1 / 3) < 1e-9 and abs(log_k[8] - 1 / 3) < 1e-12 and abs(log_k[17] + 1 / 12) < 1e-3):
        raise RuntimeError("libm: unexpected layout of the log constants")
    ltab = lbase + 18 * 8
    log_tab = []
    for i in range(_N):
        invc, logc = f64(ltab + 16 * i, 2)
        if not (0.68 < invc < 1.46 and abs(logc + math.log(invc)) < 1e-13):
            raise RuntimeError(f"libm: log table entry {i} fails its check")
        log_tab += u64(ltab + 16 * i, 2)

    # ---- __sincostab: {sin(k/128) high, low, cos(k/128) high, low} for k = 0..109 --------------
    sbase = _find_run(data, [0.0, 0.0, 1.0, 0.0, float.fromhex("0x1.fffeaaaaeeeefp-8")], "sin/cos table")
    sincos_tab = u64(sbase, 440)
    for k in range(110):
        sn, ssn, cs, ccs = f64(sbase + 32 * k, 4)
        if not (abs(sn - math.sin(k / 128)) < 2e-16 and abs(cs - math.cos(k / 128)) < 2e-16
                and abs(ssn) < 2.0 ** -53 and abs(ccs) < 2.0 ** -53):
            raise RuntimeError(f"libm: sin/cos table entry {k} fails its check")
    return {"exp_k": exp_k, "exp_tab": exp_tab, "pow_k": pow_k, "pow_tab": pow_tab,
            "log_k": log_k, "log_tab": log_tab, "sincos_tab": sincos_tab, "path": path}


def header_text(tables=None):
    """Text of ``b200_libm_tables.h`` (macros only; b200_glibc_math.cuh instantiates them)."""
    t = tables or read_tables()
    hexd = lambda v: float(v).hex()
    rows = lambda vals: ", \\\n    ".join(
        ", ".join(f"0x{v:016x}ull" for v in vals[i:i + 4]) for i in range(0, len(vals), 4))
    names_e = ["INVLN2N", "SHIFT", "NEGLN2HIN", "NEGLN2LON", "C2", "C3", "C4", "C5"]
    names_p = ["LN2HI", "LN2LO", "A0", "A1", "A2", "A3", "A4", "A5", "A6"]
    names_l = ["LN2HI", "LN2LO"] + [f"A{i}" for i in range(5)] + [f"B{i}" for i in range(11)]
    lines = [
        "// b200_libm_tables.h -- generated by brian2_b200/libm_tables.py from",
        f"// {t['path']}: the constants and tables of the host libm's exp()/pow()/log()/sin()/cos(),",
        "// so that device code reproduces the arithmetic of the oracle's own library.",
        "#pragma once",
    ]
    lines += [f"#define B200_LIBM_EXP_{n} {hexd(v)}" for n, v in zip(names_e, t["exp_k"])]
    lines += [f"#define B200_LIBM_POW_{n} {hexd(v)}" for n, v in zip(names_p, t["pow_k"])]
    lines += [f"#define B200_LIBM_LOG_{n} {hexd(v)}" for n, v in zip(names_l, t["log_k"])]
    lines.append("#define B200_LIBM_LOG_TAB \\\n    " + rows(t["log_tab"]))
    lines.append("#define B200_LIBM_SINCOS_TAB \\\n    " + rows(t["sincos_tab"]))
    lines.append("#define B200_LIBM_EXP_TAB \\\n    " + rows(t["exp_tab"]))
    lines.append("#define B200_LIBM_POW_TAB \\\n    " + rows(t["pow_tab"]))
    return "\n".join(lines) + "\n"


def write_header(directory, tables=None):
    if not host_has_fma_variants():
        raise NotImplementedError(
            "devices.b200.libm = 'glibc' restates the FMA variants of glibc's exp/expm1/pow; this "
            "host has no FMA/AVX2, its libm would run other variants")
    path = os.path.join(directory, "b200_libm_tables.h")
    text = header_text(tables)
    if not (os.path.exists(path) and open(path).read() == text):
        with open(path, "w") as f:
            f.write(text)
    return path
