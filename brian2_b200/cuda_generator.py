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

from brian2.codegen.generators.cpp_generator import CPPCodeGenerator
from brian2.codegen.permutation_analysis import (
    OrderDependenceError,
    check_for_order_independence,
)
from brian2.codegen.statements import Statement
from brian2.core.clocks import BaseClock
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
        # (EventClock -- arbitrary sample times -- is a sibling of Clock: both derive from BaseClock)
        if isinstance(owner, BaseClock) and var.name in ("t", "dt", "timestep") and isinstance(var, ArrayVariable):
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


class _CallCSE:
    """Common-subexpression elimination of stateless function calls and powers over the vector
    statements of one block.

    The state-update code Brian generates repeats whole sub-expressions textually (the
    exponential-Euler form ``_BA_x = B/A; _x = -_BA_x + (_BA_x + x)*exp(dt*A)`` contains ``A``
    twice) and the reference leaves the clean-up to the C++ compiler.  nvcc
    merges pure arithmetic, but not calls whose inlined body has control flow (``_exprel``:
    early returns around ``expm1(x)/x``; the double-double integer power; ``exp``/``expm1``
    with their range checks) -- on COBAHH these re-evaluations were half of all executed
    instructions.  Evaluating an identical expression once cannot change a single bit of the
    result, so it is done here, on the abstract code: value numbering over every ``float``
    valued stateless call / power; expressions seen more than once while their inputs are
    unchanged become ``_cse_k`` temporaries, the others are put back in place."""

    def __init__(self, variables, float_dtype, keep_exp_pow=True):
        from brian2.parsing.rendering import NodeRenderer

        self.variables = variables
        self.float_dtype = float_dtype
        self.keep_exp_pow = keep_exp_pow
        self.render = NodeRenderer().render_node

    @staticmethod
    def _is_exp_pow(node):
        import ast

        return (isinstance(node, ast.BinOp) and isinstance(node.op, ast.Pow)
                and isinstance(node.left, ast.Call) and getattr(node.left.func, "id", None) == "exp"
                and len(node.left.args) == 1)

    @staticmethod
    def _is_candidate(node):
        import ast

        if getattr(node, "dtype", None) != "float" or not getattr(node, "stateless", False):
            return False
        if getattr(node, "scalar", False):
            return False
        if isinstance(node, ast.Call):
            return True
        return isinstance(node, ast.BinOp) and isinstance(node.op, ast.Pow)

    def run(self, statements):
        import ast

        from brian2.parsing.bast import brian_ast

        available = {}      # canonical text -> temporary
        definition = {}     # temporary -> canonical text (in terms of variables and older temporaries)
        uses = {}           # temporary -> number of references
        deps = {}           # temporary -> base identifiers it is computed from
        order = []          # (position in `out`, temporary)
        out = []
        outer = self

        class Numbering(ast.NodeTransformer):
            def visit(self, node):
                if outer.keep_exp_pow and outer._is_exp_pow(node) and outer._is_candidate(node):
                    # `exp(a)**c` is ONE operation on the device (_b200_exp_pow): number the
                    # argument's sub-terms, never the inner exp() on its own
                    node.left.args[0] = self.visit(node.left.args[0])
                    node.right = self.visit(node.right)
                    return self._number(node)
                node = self.generic_visit(node)
                if outer._is_candidate(node):
                    return self._number(node)
                return node

            def _number(self, node):
                key = outer.render(node)
                name = available.get(key)
                if name is None:
                    name = f"_cse_{len(definition)}"
                    available[key] = name
                    definition[name] = key
                    uses[name] = 0
                    d = set()
                    for ident in get_identifiers(key):
                        d |= deps.get(ident, {ident})
                    deps[name] = d
                    pending.append(name)
                uses[name] += 1
                return ast.copy_location(ast.Name(id=name, ctx=ast.Load()), node)

        for stmt in statements:
            pending = []
            new_expr = None
            if not stmt.used_boolean_variables:
                try:
                    tree = Numbering().visit(brian_ast(str(stmt.expr), self.variables))
                    new_expr = self.render(tree)
                except Exception:
                    return statements          # anything unusual: leave the block alone
            for name in pending:
                order.append(name)
                out.append(("temp", name))
            out.append(("stmt", stmt, new_expr))
            # a write to `stmt.var` invalidates every temporary computed from it
            for key in [k for k, name in available.items() if stmt.var in deps[name]]:
                del available[key]

        if not any(n > 1 for n in uses.values()):
            return statements
        # single-use temporaries go back where they came from (newest first, so that nested
        # single-use terms unfold completely)
        inline = {}
        for name in reversed(order):
            if uses[name] == 1:
                inline[name] = f"({word_substitute(definition[name], inline)})"
        result = []
        for item in out:
            if item[0] == "temp":
                name = item[1]
                if name in inline:
                    continue
                result.append(Statement(name, ":=", word_substitute(definition[name], inline), "",
                                        self.float_dtype, constant=True))
            else:
                _, stmt, new_expr = item
                if new_expr is None:
                    result.append(stmt)
                    continue
                new = Statement(stmt.var, stmt.op, word_substitute(new_expr, inline), stmt.comment,
                                stmt.dtype, constant=stmt.constant, subexpression=stmt.subexpression,
                                scalar=stmt.scalar)
                new.used_boolean_variables = stmt.used_boolean_variables
                new.boolean_simplified_expressions = stmt.boolean_simplified_expressions
                result.append(new)
        return result


class CUDANodeRenderer(CPPNodeRenderer):
    """C++ expression renderer with one device-specific rewrite: ``exp(a)**c`` becomes
    ``_b200_exp_pow(a, c)`` (csrc/b200_functions.cuh) -- see there for the numerics."""

    def __init__(self, auto_vectorise=None, fuse_exp_pow=True):
        super().__init__(auto_vectorise=auto_vectorise)
        self.fuse_exp_pow = fuse_exp_pow

    # exp/expm1 -> constant-bank versions of the same algorithms (csrc/b200_functions.cuh);
    # on the host (hoisted loop-invariant scalars) they are the libm functions.  log, tanh, sinh,
    # cosh, sin, cos are CUDA's; all of them (and pow) become glibc's arithmetic with
    # prefs.devices.b200.libm = 'glibc'
    _DEVICE_MATH = {"exp": "_b200_exp", "expm1": "_b200_expm1", "log": "_b200_log",
                    "tanh": "_b200_tanh", "sinh": "_b200_sinh", "cosh": "_b200_cosh",
                    "sin": "_b200_sin", "cos": "_b200_cos"}

    def render_func(self, node):
        return self._DEVICE_MATH.get(node.id, super().render_func(node))

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
            if not getattr(var, "stateless", True):
                return False
            # functions that read arrays from their namespace (TimedArray) are evaluated on the
            # device, where those arrays live
            try:
                return not self._function_namespace_arrays(var)
            except KeyError:
                return True
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
        self._b200_scalar_write_lines = []
        written = set()
        for stmt in statements:
            ids = get_identifiers(str(stmt.expr))
            if stmt.op != ":=":
                # write to a shared (scalar) variable (`run_regularly('total += 1')`,
                # stateupdate.cpp: ALLOWS_SCALAR_WRITE): executed by ONE thread of the grid
                # after everybody's temporaries (see translate_statement_sequence)
                written.add(stmt.var)
                self._b200_scalar_write_lines.append(self.translate_statement(stmt))
                continue
            if ids & written:
                raise NotImplementedError(
                    "b200 device: a loop-invariant temporary that depends on a shared variable written "
                    f"by the same code ('{stmt.var} := {stmt.expr}')"
                )
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
        # pure scatter (`x_post += c`): the delivery loop handles several 32-synapse lines per
        # iteration, all their index loads issued ahead of the reductions (see synapses.cu)
        self._b200_unroll = 4 if statements and all(s.var in inplace_targets for s in statements) else 1
        # reads (index arrays first); in-place atomic targets are never loaded
        load_read = set(read) - inplace_targets
        lines += self.translate_to_read_arrays(load_read, write, indices)
        # The two ends of a synapse are known to the delivery loop without touching
        # `_synaptic_pre/_post[_idx]`: the spiking neuron is the CSR row, the other end comes from
        # the pathway's packed `csr_target` stream that the template loads ahead of the body.
        pathway = getattr(self.device, "_b200_current_template_kwds", {}).get("pathway")
        prepost = getattr(pathway, "prepost", None)
        if prepost in ("pre", "post"):
            ends = {"_presynaptic_idx": "_b200_src_idx" if prepost == "pre" else "_b200_tgt_idx",
                    "_postsynaptic_idx": "_b200_tgt_idx" if prepost == "pre" else "_b200_src_idx"}
            for n, line in enumerate(lines):
                m = re.match(r"^const int32_t (_presynaptic_idx|_postsynaptic_idx) = \w+\[_idx\];$", line)
                if m:
                    lines[n] = f"const int32_t {m.group(1)} = {ends[m.group(1)]};"
        # Pure scatter with per-synapse operands (`x_post += w`): the loads of the synaptic variables
        # (index `_idx`, never written here) move into the template's preload stage as well, next
        # to the index stream, so that no reduction waits for its own weight load.
        self._b200_preloads = []
        if self._b200_unroll > 1:
            kept = []
            for line in lines:
                m = re.match(r"^const (\w+) (\w+) = (\w+)\[_idx\];$", line)
                if m and m.group(2) in load_read and idx_of.get(m.group(2)) == "_idx" and m.group(2) not in write:
                    self._b200_preloads.append((m.group(1), m.group(2), m.group(3)))
                else:
                    kept.append(line)
            lines = kept
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
    # "counted" pathways: synaptic code whose effect is a function of the TARGET element only
    # ------------------------------------------------------------------------------------------
    def _countable(self, statements, read, write, indices):
        """Description of a synaptic effect that touches nothing but data of the element at the
        non-source end of the synapse (`v_post += J`, `ge_post += we`, `v_post += c*(E - v_post)`,
        with or without `(unless refractory)`): no synaptic variable, no source-side variable, no
        random numbers.  Such code gives the same result whether it runs once per event or --
        what the device does -- the events are COUNTED per target (integer reductions) and the
        owner of the target applies the statements that many times, in the reference's order
        (synapses.cpp:20-49 walks the events of one pathway one after the other).  Returns
        ``None`` if the code does not qualify, else ``{"index", "size", "read", "write"}``."""
        pathway = getattr(self.device, "_b200_current_template_kwds", {}).get("pathway")
        prepost = getattr(pathway, "prepost", None)
        if prepost not in ("pre", "post") or not statements or not write:
            return None
        tgt_index = "_postsynaptic_idx" if prepost == "pre" else "_presynaptic_idx"
        for var in self.variables.values():
            if isinstance(var, Function) and not getattr(var, "stateless", True):
                idents = set()
                for stmt in statements:
                    idents |= get_identifiers(str(stmt.expr))
                if any(self.variables.get(i) is var for i in idents):
                    return None
        arrays = set(read) | set(write)
        sizes, owners = set(), set()
        for name in arrays:
            var = self.variables[name]
            if self.variable_indices[name] != tgt_index or getattr(var, "dynamic", False):
                return None
            sizes.add(int(var.size))
            owners.add(getattr(var.owner, "name", None))
        if set(indices) - {tgt_index} or len(sizes) != 1 or len(owners) != 1:
            return None
        return {"index": tgt_index, "size": sizes.pop(), "read": sorted(read), "write": sorted(write)}

    def _counted_apply_lines(self, statements, read, write, indices, cond_write, info):
        """(loads, body, stores) of the apply loop: plain sequential code on local copies."""
        loads = self.translate_to_read_arrays(read, write, indices)
        loads = [re.sub(r"^const int32_t (_presynaptic_idx|_postsynaptic_idx) = \w+\[_idx\];$",
                        r"const int32_t \1 = _b200_tgt_idx;", line) for line in loads]
        loads += self.translate_to_declarations(read, write, indices)
        body = self.translate_to_statements(statements, cond_write)
        stores = self.translate_to_write_arrays(write)
        return loads, body, stores

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
            if prefs["devices.b200.cse"]:
                ve_block = _CallCSE(
                    self.variables, prefs["core.default_float_dtype"],
                    keep_exp_pow=bool(prefs["devices.b200.fuse_exp_pow"]),
                ).run(ve_block)
            sc_read, sc_write, sc_indices, sc_cond = self.arrays_helper(sc_block)
            ve_read, ve_write, ve_indices, ve_cond = self.arrays_helper(ve_block)
            if sc_write:
                # stateupdate.cpp is ALLOWS_SCALAR_WRITE (`run_regularly('shared_var += ...')`).  Every
                # thread evaluates the scalar block, so the write itself is done by one elected
                # thread; the per-element code of the SAME code object must not read the variable
                # (it could see either value) -- other code objects are ordered by grid barriers
                # (the barrier analysis sees a shared write).
                clash = sorted(set(sc_write) & (set(ve_read) | set(ve_write)))
                if clash or any(c is not None for c in sc_cond.values()):
                    raise NotImplementedError(
                        "b200 device: per-element code that reads a shared (scalar) variable written "
                        f"by the same code object ({', '.join(clash)} in '{self.name}')"
                    )
            if self.template_name == "synapses_create_generator":
                # the device version of connect() evaluates index arithmetic and rand() only;
                # conditions that read state variables of the connected groups need the
                # reference's host path (prefs.devices.b200.construction = 'reference')
                # ... and the index arrays `i` of the connected groups (arange arrays: the value
                # IS the index), which subgroup sources / targets bring in
                identity = getattr(self, "_b200_identity_arrays", {})
                aranges = {v: start for v, _, start in self.device.arange_arrays}
                for name in sorted(ve_read | sc_read):
                    var = self.variables[name]
                    start = aranges.get(var) if name not in ve_write else None
                    if start is not None:
                        identity[self.get_array_name(var)] = int(start)
                self._b200_identity_arrays = identity
                touched = sorted(n for n in (ve_read | ve_write | sc_read | ve_indices | sc_indices)
                                 if self.get_array_name(self.variables[n]) not in identity)
                if touched:
                    raise NotImplementedError(
                        "b200 sharded construction: connect() expressions that read arrays "
                        f"({', '.join(touched)}) are not supported on the device"
                    )
            # scalar variables needed by the vector code are read once, in the scalar block
            for varname in set(ve_read):
                var = self.variables[varname]
                if var.scalar and varname not in ve_write:
                    sc_read.add(varname)
                    ve_read.remove(varname)

            # ---- scalar block
            read_lines = self.translate_to_read_arrays(sc_read, sc_write, sc_indices)
            if sc_write:
                read_lines += self.translate_to_declarations(sc_read, sc_write, sc_indices)
            hoisted, host_lines, dev_lines = self._split_scalar_block(sc_block, sc_read)
            if sc_write:
                dev_lines += (["if (_ctx.gbid == 0 && threadIdx.x == 0)", "{"]
                              + ["    " + ln for ln in self._b200_scalar_write_lines]
                              + ["    " + ln for ln in self.translate_to_write_arrays(sc_write)]
                              + ["}"])
            sc_code[block_name] = stripped_deindented_lines("\n".join(read_lines + dev_lines))
            scal_members[block_name] = hoisted
            fill = [f"_sc.{name} = {name};" for _, name in hoisted]
            # host flavour reads only invariant scalars (dt, constants): `t` reads are harmless
            scal_host[block_name] = stripped_deindented_lines(
                "\n".join(read_lines + host_lines + fill)
            )

            # ---- vector block
            counted, dual = None, False
            if self._is_synaptic_effect() and len(ve_block) and prefs["devices.b200.counted_pathways"] != "never":
                counted = self._countable(ve_block, ve_read, ve_write, ve_indices)
                if counted is not None and prefs["devices.b200.counted_pathways"] == "auto":
                    # Pure `x_post += constant` scatters: sparse rows are served best by fp
                    # reductions (one phase less per step), so that stays the delivery -- "dual":
                    # the counted form is generated next to it and only used when the rows turn
                    # out to be dense at run time (owner-computes over target tiles).  Code that
                    # reads target-side data (`(unless refractory)`, `v_post += c*(E - v_post)`)
                    # is always counted.
                    dual = not (set(ve_read) - set(ve_write)) and all(
                        s.inplace and s.op in ("+=", "-=") and not (get_identifiers(str(s.expr)) & set(ve_write))
                        for s in ve_block)
            if counted is not None:
                loads, body, stores = self._counted_apply_lines(ve_block, ve_read, ve_write, ve_indices, ve_cond, counted)
                kwds["b200_apply_loads"] = stripped_deindented_lines("\n".join(loads))
                kwds["b200_apply_body"] = stripped_deindented_lines("\n".join(body))
                kwds["b200_apply_stores"] = stripped_deindented_lines("\n".join(stores))
                name_of = lambda n: self.device.get_array_name(self.variables[n], access_data=False)
                access["counted"] = {"size": counted["size"], "dual": dual,
                                     "read": sorted(name_of(n) for n in counted["read"]),
                                     "write": sorted(name_of(n) for n in counted["write"])}
            if counted is not None and not dual:
                ve_code[block_name] = ["atomicAdd(_b200_hits + _b200_tgt_idx, 1);"]
                self._b200_unroll, self._b200_preloads = 4, []
                for name in sc_read | sc_indices:
                    var = self.variables.get(name)
                    if isinstance(var, ArrayVariable):
                        access["read"].add(self.device.get_array_name(var, access_data=False))
                continue
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
                    if name in sc_write:     # one thread writes what every thread of the grid may read
                        key = "scattered_write"
                        access.setdefault("scalar_write", set()).add(
                            self.device.get_array_name(var, access_data=False))
                    access[key].add(self.device.get_array_name(var, access_data=False))

        if set(scal_host.keys()) == {None}:
            scal_host, scal_members = scal_host[None], scal_members[None]
        kwds["b200_scalar_host"] = scal_host
        kwds["b200_scalar_members"] = scal_members
        kwds["b200_identity_arrays"] = sorted(getattr(self, "_b200_identity_arrays", {}).items())
        kwds["b200_counted"] = access.get("counted")
        kwds["b200_serial"] = serial
        kwds["b200_unroll"] = 1 if serial else getattr(self, "_b200_unroll", 1)
        kwds["b200_gather_unroll"] = 4 if kwds["b200_unroll"] > 1 else 1
        kwds["b200_preloads"] = [] if serial else list(getattr(self, "_b200_preloads", []))
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
        if var.constant and var.read_only and not var.scalar:
            # never written inside a run (e.g. `_synaptic_pre/_post`, synapses.py:1386-1391):
            # non-coherent loads are safe and the compiler may batch them ahead of the atomics
            return f"const {ctype}* __restrict__ {pointer_name} = _A.{array_name};"
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
                    # Arrays in a function's namespace (e.g. the values of a TimedArray): the C++
                    # target keeps them in file-scope statics assigned inside the code object
                    # (cpp_generator.py:459-474).  Device code reads them through the array
                    # table instead: `_namespace<key>` becomes a macro for `_A.<key>` and the
                    # device registers <key> for upload next to the ordinary arrays.
                    ns_keys = self._function_namespace_arrays(variable)
                    for key, (ctype, size) in ns_keys.items():
                        self.device._b200_func_arrays[key] = (ctype, size)
                        support_code.append(f"#define _namespace{key} (_A.{key})")
                    for code in sc:
                        if any(code.strip() == f"static {ctype}* _namespace{key};"
                               for key, (ctype, _) in ns_keys.items()):
                            continue
                        support_code.append(self._device_support_code(code))
                    ps = [line for line in ps
                          if not any(line.strip().startswith(f"_namespace{key} =") for key in ns_keys)]
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

    def _function_namespace_arrays(self, variable, seen=None):
        """{namespace key: (ctype, size)} of the arrays a function (and its dependencies) brings
        along in its namespace."""
        seen = seen if seen is not None else set()
        out = {}
        if variable in seen:
            return out
        seen.add(variable)
        impl = variable.implementations[self.codeobj_class]
        for key, value in (impl.get_namespace(self.owner) or {}).items():
            if hasattr(value, "dtype") and getattr(value, "shape", ()) != ():
                out[key] = (self.c_data_type(value.dtype), int(value.size))
        for dep in (impl.dependencies or {}).values():
            if isinstance(dep, Function):
                out.update(self._function_namespace_arrays(dep, seen))
        return out

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
