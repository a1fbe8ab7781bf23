"""Code object classes of the ``b200`` device.

* `B200CodeObject` -- in-loop code objects, rendered from the CUDA templates in
  ``brian2_b200/templates`` by `CUDACodeGenerator`.
* `B200HostCodeObject` -- run-once code objects (variable initialisation, ``Synapses.connect``):
  these stay *host C++ exactly as the reference emits them* (same templates, same
  ``RandomGenerator``/mt19937 stream, ``templates/objects.cpp:426-479``), so connectivity and
  initial state are bit-identical to ``cpp_standalone`` for the same ``seed()``.

Pattern: ``brian2/devices/cpp_standalone/codeobject.py:85-172``.
"""
from brian2.codegen.codeobject import constant_or_scalar as host_constant_or_scalar
from brian2.codegen.generators.cpp_generator import CPPCodeGenerator, c_data_type
from brian2.codegen.targets import codegen_targets
from brian2.codegen.templates import Templater
from brian2.core.functions import DEFAULT_FUNCTIONS
from brian2.core.preferences import prefs
from brian2.devices.cpp_standalone.codeobject import CPPStandaloneCodeObject
from brian2.devices.device import get_device

from .cuda_generator import CUDACodeGenerator

__all__ = ["B200CodeObject", "B200HostCodeObject", "B200ConnectCodeObject", "B200ShardHostCodeObject",
           "DEVICE_TEMPLATES", "HOST_TEMPLATES", "RUN_ONCE_DEVICE_TEMPLATES"]

#: templates that run inside the time loop and have a CUDA version
DEVICE_TEMPLATES = {
    "stateupdate",
    "threshold",
    "reset",
    "synapses",
    "synapses_push_spikes",
    "spikemonitor",
    "statemonitor",
    "ratemonitor",
    "spikegenerator",
    "summed_variable",
}

#: run-once templates that have a device version (used under "sharded construction" only)
RUN_ONCE_DEVICE_TEMPLATES = {"synapses_create_generator"}

#: templates that only ever run once, on the host, before/between runs
HOST_TEMPLATES = {
    "group_variable_set",
    "group_variable_set_conditional",
    "synapses_create_generator",
    "synapses_create_array",
}


def device_constant_or_scalar(varname, variable):
    """Device flavour of ``codegen/codeobject.py:32``: scalar arrays are read through the
    local device pointer of the code object."""
    if variable.array:
        return f"_ptr{get_device().get_array_name(variable)}[0]"
    return f"{varname}"


def b200_field(var):
    """Name of the field of the device array table ``_A`` that holds ``var``."""
    return get_device().arrays[var]


class _DualTemplater(Templater):
    """CUDA templates of this package first, the reference's C++ templates as fallback.
    (`Templater.derive`, codegen/templates.py:152, cannot mix file extensions.)  The fallback is
    only used to *look up* run-once templates -- `B200Device.code_object` re-routes those to
    `B200HostCodeObject`."""

    def __getattr__(self, item):
        try:
            return self.templates.get_template(item)
        except KeyError:
            try:
                return getattr(CPPStandaloneCodeObject.templater, item)
            except AttributeError as ex:
                raise AttributeError(item) from ex


class B200CodeObject(CPPStandaloneCodeObject):
    """In-loop code object: a ``__device__`` function + kernel for sm_100a."""

    templater = _DualTemplater(
        "brian2_b200",
        ".cu",
        env_globals={
            "c_data_type": c_data_type,
            "constant_or_scalar": device_constant_or_scalar,
            "b200_host_constant_or_scalar": host_constant_or_scalar,
            "b200_field": b200_field,
            "prefs": prefs,
            "zip": zip,
        },
    )
    generator_class = CUDACodeGenerator


class B200HostCodeObject(CPPStandaloneCodeObject):
    """Run-once host code object: the reference's own C++ templates and generator."""

    templater = CPPStandaloneCodeObject.templater
    generator_class = CPPCodeGenerator


class B200ConnectCodeObject(B200CodeObject):
    """``Synapses.connect`` on the device (templates/synapses_create_generator.cu): a run-once code
    object whose kernel is compiled into the project's CUDA translation unit."""


class B200ShardHostCodeObject(B200HostCodeObject):
    """Run-once host code object that initialises the variables of a `Synapses` object whose
    connectivity was created per rank ("sharded construction"): same reference templates, but
    ``rand()``/``randn()`` are pure functions of the synapse (csrc/b200_synrng.h) instead of
    draws from the sequential host stream."""


codegen_targets.add(B200CodeObject)
codegen_targets.add(B200HostCodeObject)
codegen_targets.add(B200ConnectCodeObject)
codegen_targets.add(B200ShardHostCodeObject)

# ---------------------------------------------------------------------------------------------
# Function implementations for device code.  Everything that is a plain libm call is inherited
# from the C++ generator through the MRO lookup (core/functions.py:352-390); helpers that the
# reference pastes as host-only support code (cpp_generator.py:603-656) live in
# csrc/b200_functions.cuh instead and only need their name registered here.
# ---------------------------------------------------------------------------------------------
for _func, _name in [
    ("exprel", "_exprel"),
    ("abs", "_brian_abs"),
    ("clip", "_clip"),
    ("sign", "_sign"),
    ("timestep", "_timestep"),
    ("int", "_b200_int"),
]:
    DEFAULT_FUNCTIONS[_func].implementations.add_implementation(
        CUDACodeGenerator, code=None, name=_name
    )

# In-loop random numbers: counter-based Philox streams (csrc/b200_runtime.cuh); the templates
# create the per-element generator `_rng`.
DEFAULT_FUNCTIONS["rand"].implementations.add_implementation(
    B200CodeObject,
    code={"support_code": "", "hashdefine_code": "#define _rand(_i) b200::rng_uniform(_rng)"},
    name="_rand",
)
DEFAULT_FUNCTIONS["randn"].implementations.add_implementation(
    B200CodeObject,
    code={"support_code": "", "hashdefine_code": "#define _randn(_i) b200::rng_normal(_rng)"},
    name="_randn",
)


# poisson(lam): the reference's samplers on the element's Philox stream (b200::rng_poisson)
DEFAULT_FUNCTIONS["poisson"].implementations.add_implementation(
    B200CodeObject,
    code={"support_code": "", "hashdefine_code": "#define _poisson(_lam, _i) b200::rng_poisson(_rng, (_lam))"},
    name="_poisson",
)


def _synapses_of(owner):
    """The `Synapses` object behind ``owner`` (itself, or the owner of a `SynapticPathway`)."""
    return getattr(owner, "synapses", owner)


def _generate_synapse_rng_code(func, owner):
    import zlib

    S = _synapses_of(owner)
    device = get_device()
    pre = device.get_array_name(S.variables["_synaptic_pre"], access_data=False)
    post = device.get_array_name(S.variables["_synaptic_post"], access_data=False)
    method = {"rand": "uniform", "randn": "normal"}[func]
    stream = zlib.crc32(f"{owner.name}.{func}".encode())
    code = f"""
    static b200::SynapseRng _b200_synrng_{func}({stream}u);
    inline double _{func}(const int _vectorisation_idx) {{
        return _b200_synrng_{func}.{method}(_vectorisation_idx, brian::{pre}, brian::{post});
    }}
    """
    return {"support_code": code}


for _func in ("rand", "randn"):
    DEFAULT_FUNCTIONS[_func].implementations.add_dynamic_implementation(
        B200ShardHostCodeObject,
        code=(lambda f: (lambda owner: _generate_synapse_rng_code(f, owner)))(_func),
        namespace=lambda owner: {},
        name=f"_{_func}",
    )


# ---------------------------------------------------------------------------------------------
# BinomialFunction (PoissonInput): the reference generates one sampler per instance through the
# class-level table ``BinomialFunction.implementations`` (input/binomial.py:168, applied at
# :207-220) -- "this container can be extended by other code generation targets".  Device
# flavour: the same inversion / normal-approximation algorithm (input/binomial.py:85-143, itself
# numpy's rk_binomial_inversion), drawing from the element's Philox stream `_rng` that the
# templates create; the call `name(_vectorisation_idx)` is a macro that passes `_rng` along.
# ---------------------------------------------------------------------------------------------
def _generate_cuda_binomial(n, p, use_normal, name):
    from brian2.input.binomial import _pre_calc_constants, _pre_calc_constants_approximated

    if use_normal:
        loc, scale = _pre_calc_constants_approximated(n, p)
        # (the C++ target returns `float` here, input/binomial.py:91: same rounding kept)
        support = f"""
        __device__ __forceinline__ float {name}_b200(b200::Rng& _rng)
        {{
            return b200::rng_normal(_rng) * {scale:.15f} + {loc:.15f};
        }}
        """
    else:
        reverse, q, P, qn, bound = _pre_calc_constants(n, p)
        ret = f"{int(n)}-X" if reverse else "X"
        support = f"""
        __device__ __forceinline__ long {name}_b200(b200::Rng& _rng)
        {{
            long X = 0;
            double px = {qn:.15f};
            double U = b200::rng_uniform(_rng);
            while (U > px)
            {{
                X++;
                if (X > {bound:.15f})
                {{
                    X = 0;
                    px = {qn:.15f};
                    U = b200::rng_uniform(_rng);
                }} else
                {{
                    U -= px;
                    px = (({int(n)}-X+1) * {P:.15f} * px)/(X*{q:.15f});
                }}
            }}
            return {ret};
        }}
        """
    code = {"support_code": support, "hashdefine_code": f"#define {name}(_i) {name}_b200(_rng)"}
    return code, {}


def _register_binomial():
    from brian2.input.binomial import BinomialFunction

    BinomialFunction.implementations["cuda"] = _generate_cuda_binomial


_register_binomial()
