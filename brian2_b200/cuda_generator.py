"""CUDA code generator for the ``b200`` device.

Turns Brian's abstract code (already split into scalar / vector statements by
``brian2/codegen/translation.py:181`` ``make_statements``) into the bodies of sm_100a
``__device__`` functions.  It derives from the reference's ``CPPCodeGenerator``
(``brian2/codegen/generators/cpp_generator.py:200``) for expression rendering and function
lookup, and changes what a GPU needs changed:

* **Scalar (loop-invariant) block** (``stateupdate.cpp:7-9``): every ``_lio_k`` whose value cannot
  change during a run (depends only on constants, ``dt`` and stateless functions) is *hoisted to
  the host*: it is computed once per run with the host's libm -- bit-identical to what the
  reference's generated C++ computes -- and handed to the kernels in a small struct.  Statements
  that depend on ``t`` or on mutable shared variables stay on the device.
* **Array access**: pointers come from the ``__constant__`` array table ``_A``; clock variables
  come from the by-value clock struct; event spaces resolve to the current slot of the device
  spike ring.
* **Synaptic code** (``templates/synapses.cpp:11-50``): the C++ target emits load / modify /
  store for ``x_post += w`` which is only correct sequentially (and is why the reference runs it
  under ``#pragma omp master``, ``synapses.cpp:25-36``).  If
  ``check_for_order_independence`` (``codegen/permutation_analysis.py:16``) passes, writes to
  variables that are not indexed by the synapse index become in-place atomics
  (``+=``/``-=`` -> ``atomicAdd``, ``*=``/``/=`` -> CAS loop, ``=`` -> plain store); otherwise the
  template falls back to a serial walk in the reference's delivery order.
"""
import re

from brian2.codegen.generators.cpp_generator import CPPCodeGenerator, c_data_type
from brian2.codegen.permutation_analysis import (
    OrderDependenceError,
    check_for_order_independence,
)
from brian2.codegen.statements import Statement
from brian2.core.clocks import Clock
from brian2.core.preferences import prefs
from brian2.parsing.rendering import CPPNodeRenderer
from brian2.core.functions import Function
from brian2.core.variables import ArrayVariable, Constant
from brian2.utils.stringtools import (
    deindent,
    get_identifiers,
    stripped_deindented_lines,
    word_substitute,
)

__all__ = ["CUDACodeGenerator", "is_eventspace", "clock_field"]

#: functions whose implementation lives in csrc/b200_functions.cuh -- their C++ support code
#: (host only) must not be pasted into device code
_BUILTIN_DEVICE_FUNCTIONS = {
    "_timestep", "_exprel", "_clip", "_sign", "_brian_abs", "int_", "_b200_int",
    "_brian_mod", "_brian_floordiv", "_brian_pow", "_b200_exp_pow",
}


def is_eventspace(var):
    """``True`` for the ``_spikespace``-like variables (int32, N+1 entries; neurongroup.py:798)."""
    name = getattr(var, "name", "")
    return (
        isinstance(var, ArrayVariable)
        and name.startswith("_")
        and name.endswith("space")
        and not getattr(var, "dynamic", False)
    )


def clock_field(var):
    """If ``var`` is the ``t``/``dt``/``timestep`` array of a `Clock`, return the field name."""
    owner = getattr(var, "owner", None)
    try:
        if isinstance(owner, Clock) and var.name in ("t", "dt", "timestep") and isinstance(var, ArrayVariable):
            return var.name
    except ReferenceError:
        pass
    return None


def _annotate_host_device(code):
    """Best effort: make user supplied C++ support code callable from device code."""
    out = []
    pattern = re.compile(
        r"^(\s*)((?:static\s+|inline\s+)*)((?:const\s+)?(?:unsigned\s+)?[A-Za-z_][\w:<>]*(?:\s*[\*&])?\s+[A-Za-z_]\w*\s*\([^;{]*\)\s*\{?\s*)$"
    )
    for line in code.split("\n"):
        m = pattern.match(line)
        if m and "return" not in line and not line.strip().startswith(("if", "for", "while", "else", "#")):
            line = f"{m.group(1)}__host__ __device__ {m.group(2)}{m.group(3)}"
        out.append(line)
    return "\n".join(out)


class CUDANodeRenderer(CPPNodeRenderer):
    """C++ expression renderer with one device-specific rewrite: ``exp(a)**c`` becomes
    ``_b200_exp_pow(a, c)`` (csrc/b200_functions.cuh) -- see there for the numerics."""

    def __init__(self, auto_vectorise=None, fuse_exp_pow=True):
        super().__init__(auto_vectorise=auto_vectorise)
        self.fuse_exp_pow = fuse_exp_pow

    def render_BinOp(self, node):
        if (
            self.fuse_exp_pow
            and node.op.__class__.__name__ == "Pow"
            and node.left.__class__.__name__ == "Call"
            and getattr(node.left.func, "id", None) == "exp"
            and len(node.left.args) == 1
        ):
            return (
                f"_b200_exp_pow({self.render_node(node.left.args[0])}, "
                f"{self.render_node(node.right)})"
            )
        return super().render_BinOp(node)


class CUDACodeGenerator(CPPCodeGenerator):
    """CUDA (sm_100a) language for the in-loop code objects of the ``b200`` device."""

    class_name = "cuda"

    # the helper functions come from b200_functions.cuh instead of pasted templates
    universal_support_code = ""

    # ------------------------------------------------------------------------------------------
    @property
    def restrict(self):
        # No __restrict__ on state arrays: inside the persistent kernel the same array is read by
        # one code object and written by another, and ld.global.nc on such data would be stale.
        return " "

    def translate_expression(self, expr):
        expr = word_substitute(expr, self.func_name_replacements)
        return (
            CUDANodeRenderer(
                auto_vectorise=self.auto_vectorise,
                fuse_exp_pow=bool(prefs["devices.b200.fuse_exp_pow"]),
            )
            .render_expr(expr)
            .strip()
        )

    def _is_synaptic_effect(self):
        return self.template_name == "synapses"

    # ------------------------------------------------------------------------------------------
    # scalar block: split into host-hoisted and device-evaluated statements
    # ------------------------------------------------------------------------------------------
    def _invariant_name(self, name, known):
        if name in known:
            return known[name]
        var = self.variables.get(name)
        if var is None:
            return name in ("True", "False", "inf", "nan", "true", "false")
        if isinstance(var, Constant):
            return True
        if isinstance(var, Function):
            return bool(getattr(var, "stateless", True))
        if isinstance(var, ArrayVariable):
            if not (var.scalar or self.variable_indices[name] == "0"):
                return False
            if clock_field(var) == "dt":
                return True
            return bool(var.constant)
        return False

    def _split_scalar_block(self, statements, read):
        """Return (hoisted [(ctype, name)], host_lines, dev_lines) for one scalar block."""
        known = {}
        hoisted, host_lines, dev_lines = [], [], []
        for stmt in statements:
            if stmt.op != ":=":
                raise NotImplementedError(
                    "b200 device: writes to shared (scalar) variables inside the simulation loop "
                    f"are not supported yet (statement '{stmt.var} {stmt.op} ...')"
                )
            ids = get_identifiers(str(stmt.expr))
            inv = all(self._invariant_name(i, known) for i in ids)
            known[stmt.var] = inv
            line = self.translate_statement(stmt)
            if inv:
                ctype = self.c_data_type(stmt.dtype)
                hoisted.append((ctype, stmt.var))
                host_lines.append(line)
                dev_lines.append(f"const {ctype} {stmt.var} = _sc.{stmt.var};")
            else:
                dev_lines.append(line)
        return hoisted, host_lines, dev_lines

    # ------------------------------------------------------------------------------------------
    # vector block of a synaptic pathway: atomics for non-synaptic targets
    # ------------------------------------------------------------------------------------------
    def _synaptic_vector_lines(self, statements, read, write, indices, cond_write):
        idx_of = self.variable_indices
        nonmain_written = {
            v for v in write if idx_of[v] not in ("_idx", "0")
        }
        inplace_targets = {
            s.var for s in statements if s.var in nonmain_written and s.inplace
        }
        lines = []
        # reads (index arrays first); in-place atomic targets are never loaded
        load_read = set(read) - inplace_targets
        lines += self.translate_to_read_arrays(load_read, write, indices)
        lines += self.translate_to_declarations(load_read | inplace_targets, write, indices)
        tmp_count = 0
        for stmt in statements:
            condvar = cond_write.get(stmt.var)
            stmt_lines = []
            if stmt.var in inplace_targets:
                var = self.variables[stmt.var]
                ctype = self.c_data_type(var.dtype)
                tmp = f"_b200_upd_{tmp_count}"
                tmp_count += 1
                proxy = Statement(tmp, "=", stmt.expr, stmt.comment, stmt.dtype)
                proxy.used_boolean_variables = stmt.used_boolean_variables
                proxy.boolean_simplified_expressions = stmt.boolean_simplified_expressions
                stmt_lines.append(f"{ctype} {tmp};")
                stmt_lines.append(self.translate_statement(proxy))
                target = f"&{self.get_array_name(var)}[{idx_of[stmt.var]}]"
                if stmt.op == "+=":
                    stmt_lines.append(f"b200::atomic_add({target}, {tmp});")
                elif stmt.op == "-=":
                    stmt_lines.append(f"b200::atomic_add({target}, -{tmp});")
                elif stmt.op == "*=":
                    stmt_lines.append(f"b200::atomic_mul({target}, {tmp});")
                elif stmt.op == "/=":
                    stmt_lines.append(f"b200::atomic_div({target}, {tmp});")
                else:
                    raise NotImplementedError(
                        f"b200 device: in-place operator '{stmt.op}' on a non-synaptic variable"
                    )
            else:
                stmt_lines.append(self.translate_statement(stmt))
                if stmt.var in nonmain_written:
                    var = self.variables[stmt.var]
                    stmt_lines.append(
                        f"{self.get_array_name(var)}[{idx_of[stmt.var]}] = {stmt.var};"
                    )
            if condvar is not None:
                lines.append(f"if({condvar})")
                lines.append("{")
                lines += ["    " + ln for ln in "\n".join(stmt_lines).split("\n")]
                lines.append("}")
            else:
                lines += stmt_lines
        lines += self.translate_to_write_arrays(set(write) - nonmain_written)
        return lines

    # ------------------------------------------------------------------------------------------
    def translate_statement_sequence(self, sc_statements, ve_statements):
        assert set(sc_statements.keys()) == set(ve_statements.keys())
        kwds = self.determine_keywords()
        sc_code, ve_code = {}, {}
        scal_host, scal_members = {}, {}
        serial = False
        access = {"read": set(), "write": set(), "scattered_write": set(), "scattered_read": set()}

        for block_name in sc_statements:
            sc_block = sc_statements[block_name]
            ve_block = ve_statements[block_name]
            sc_read, sc_write, sc_indices, sc_cond = self.arrays_helper(sc_block)
            ve_read, ve_write, ve_indices, ve_cond = self.arrays_helper(ve_block)
            # scalar variables needed by the vector code are read once, in the scalar block
            for varname in set(ve_read):
                var = self.variables[varname]
                if var.scalar and varname not in ve_write:
                    sc_read.add(varname)
                    ve_read.remove(varname)

            # ---- scalar block
            read_lines = self.translate_to_read_arrays(sc_read, sc_write, sc_indices)
            hoisted, host_lines, dev_lines = self._split_scalar_block(sc_block, sc_read)
            sc_code[block_name] = stripped_deindented_lines("\n".join(read_lines + dev_lines))
            scal_members[block_name] = hoisted
            fill = [f"_sc.{name} = {name};" for _, name in hoisted]
            # host flavour reads only invariant scalars (dt, constants): `t` reads are harmless
            scal_host[block_name] = stripped_deindented_lines(
                "\n".join(read_lines + host_lines + fill)
            )

            # ---- vector block
            if self._is_synaptic_effect() and len(ve_block):
                try:
                    if self.has_repeated_indices(ve_block):
                        check_for_order_independence(
                            ve_block, self.variables, self.variable_indices
                        )
                except OrderDependenceError:
                    serial = True
            if self._is_synaptic_effect() and not serial:
                lines = self._synaptic_vector_lines(ve_block, ve_read, ve_write, ve_indices, ve_cond)
            else:
                lines = []
                lines += self.translate_to_read_arrays(ve_read, ve_write, ve_indices)
                lines += self.translate_to_declarations(ve_read, ve_write, ve_indices)
                lines += self.translate_to_statements(ve_block, ve_cond)
                lines += self.translate_to_write_arrays(ve_write)
            ve_code[block_name] = stripped_deindented_lines("\n".join(lines))

            # ---- access summary for the barrier analysis of the persistent kernel
            for name in ve_read | sc_read | ve_indices | sc_indices:
                var = self.variables.get(name)
                if isinstance(var, ArrayVariable):
                    key = "read" if self.variable_indices[name] in ("_idx", "0") else "scattered_read"
                    access[key].add(self.device.get_array_name(var, access_data=False))
            for name in ve_write | sc_write:
                var = self.variables.get(name)
                if isinstance(var, ArrayVariable):
                    key = "write" if self.variable_indices[name] in ("_idx", "0") else "scattered_write"
                    access[key].add(self.device.get_array_name(var, access_data=False))

        if set(scal_host.keys()) == {None}:
            scal_host, scal_members = scal_host[None], scal_members[None]
        kwds["b200_scalar_host"] = scal_host
        kwds["b200_scalar_members"] = scal_members
        kwds["b200_serial"] = serial
        # remembered by the device for the barrier analysis of the persistent kernel
        access["serial"] = serial
        self.device._b200_access[self.name] = access
        return sc_code, ve_code, kwds

    # ------------------------------------------------------------------------------------------
    # keywords: pointers, support code
    # ------------------------------------------------------------------------------------------
    def _device_pointer_line(self, var):
        array_name = self.device.get_array_name(var)
        pointer_name = self.get_array_name(var)
        ctype = self.c_data_type(var.dtype)
        field = clock_field(var)
        if field is not None:
            return f"const {ctype}* {pointer_name} = &_clks.{var.owner.name}.{field};"
        if is_eventspace(var):
            # generic access = the compacted list of the current step (reference layout); the
            # device templates themselves read the segments through b200::view_* instead
            clk = var.owner.clock.name
            return (
                f"const {ctype}* {pointer_name} = b200::compact_slot(_A._es{array_name}, "
                f"_clks.{clk}.timestep);"
            )
        return f"{ctype}* {self.restrict}{pointer_name} = _A.{array_name};"

    def _host_pointer_line(self, var):
        array_name = self.device.get_array_name(var)
        pointer_name = self.get_array_name(var)
        ctype = self.c_data_type(var.dtype)
        return f"{ctype}* {pointer_name} = {array_name};"

    def determine_keywords(self):
        pointers, host_pointers = [], []
        handled = set()
        for var in self.variables.values():
            if isinstance(var, ArrayVariable):
                pointer_name = self.get_array_name(var)
                if pointer_name in handled:
                    continue
                if getattr(var, "ndim", 1) > 1:
                    continue
                handled.add(pointer_name)
                pointers.append(self._device_pointer_line(var))
                host_pointers.append(self._host_pointer_line(var))

        user_functions, support_code, hash_defines = [], [], []
        added = set()
        uses_rng = False
        for varname, variable in list(self.variables.items()):
            if isinstance(variable, Function):
                if not getattr(variable, "stateless", True):
                    uses_rng = True
                user_func = self._add_user_function(varname, variable, added)
                if user_func is not None:
                    hd, ps, sc, uf = user_func
                    user_functions.extend(uf)
                    for code in sc:
                        support_code.append(self._device_support_code(code))
                    pointers.extend(ps)
                    host_pointers.extend(ps)
                    hash_defines.extend(hd)

        return {
            "pointers_lines": stripped_deindented_lines("\n".join(pointers)),
            "host_pointers_lines": stripped_deindented_lines("\n".join(host_pointers)),
            "support_code_lines": stripped_deindented_lines("\n".join(support_code)),
            "hashdefine_lines": stripped_deindented_lines("\n".join(hash_defines)),
            "denormals_code_lines": [],
            "b200_uses_rng": uses_rng,
        }

    def _device_support_code(self, code):
        code = deindent(code)
        if not code.strip():
            return ""
        if "__device__" in code or "B200_HD" in code:
            return code
        # pasted C++ of a function we provide natively -> drop
        names = set(re.findall(r"\b(_[A-Za-z_]\w*)\s*\(", code))
        if names and names <= _BUILTIN_DEVICE_FUNCTIONS:
            return ""
        if code.strip().startswith("#define"):
            return code
        return _annotate_host_device(code)
