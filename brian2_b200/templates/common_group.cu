{# Common skeleton of every in-loop (device) code object of the b200 device.
   Counterpart of brian2/devices/cpp_standalone/templates/common_group.cpp, but the main block is
   a __device__ function that is inlined both into a per-code-object kernel (profiling /
   multi-clock mode) and into the persistent step kernel. #}
{% macro cpp_file() %}
{% block b200_file %}
// ===== code object {{codeobj_name}} (template: {{b200_template_name}}) =====
namespace _co_{{codeobj_name}} {
    ////// SUPPORT CODE ///////
    {{support_code_lines|autoindent}}
    // loop-invariant scalars, computed once per run on the host (bit-identical libm)
    struct Scal {
        {% for ctype, name in b200_scalar_members %}
        {{ctype}} {{name}};
        {% endfor %}
        int _unused;
    };
}

////// HASH DEFINES ///////
{{hashdefine_lines|autoindent}}

static void _hostscal_{{codeobj_name}}(_co_{{codeobj_name}}::Scal& _sc)
{
    using namespace brian;
    using namespace _co_{{codeobj_name}};
    const size_t _vectorisation_idx = -1;
    ///// CONSTANTS ///////////
    %CONSTANTS%
    ///// POINTERS ////////////
    {{host_pointers_lines|autoindent}}
    {{b200_scalar_host|autoindent}}
    _sc._unused = 0;
}

__device__ __forceinline__ void _dev_{{codeobj_name}}(const b200::Ctx& _ctx, const _B200Clocks& _clks,
                                                     const _co_{{codeobj_name}}::Scal& _sc)
{
    using namespace _co_{{codeobj_name}};
    ///// CONSTANTS ///////////
    %CONSTANTS_DEV%
    ///// POINTERS ////////////
    {{pointers_lines|autoindent}}
    {% block maincode %}
    {% endblock %}
}

__global__ void __launch_bounds__(b200::kBlock, {{prefs.devices.b200.ctas_per_sm}})
_kernel_{{codeobj_name}}(const _B200Clocks _clks, const _co_{{codeobj_name}}::Scal _sc)
{
    const b200::Ctx _ctx{(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.x, (int)gridDim.x, _A._rank, _A._world};
    b200::view_reset();
    _dev_{{codeobj_name}}(_ctx, _clks, _sc);
}

B200_REGISTER_KERNEL(_kernel_{{codeobj_name}})

void _run_{{codeobj_name}}()
{
    // Called from main() outside of a Network::run (e.g. StateMonitor.record_single_timestep(),
    // monitors/statemonitor.py:398-420): the host mirrors are the truth between runs, so the
    // single launch is bracketed by the same upload / download as a run.
    const bool _standalone_call = !Network::_globally_running;
    if (_standalone_call) { _b200_upload(); _b200_prepare_steps(1, true); }
    _co_{{codeobj_name}}::Scal _sc;
    _hostscal_{{codeobj_name}}(_sc);
    {% block host_prelaunch %}
    {% endblock %}
    _b200_launch_begin("{{codeobj_name}}");
    _kernel_{{codeobj_name}}<<<_b200_grid_size(), b200::kBlock, 0, b200::state().stream>>>(_b200_clocks_now(), _sc);
    {% block host_postlaunch %}
    {% endblock %}
    _b200_launch_end("{{codeobj_name}}");
    if (_standalone_call) { B200_CUDA(cudaStreamSynchronize(b200::state().stream)); _b200_download(); }
}
{% block extra_device_code %}
{% endblock %}
{% endblock %}
{% endmacro %}


{% macro h_file() %}
#ifndef _INCLUDED_{{codeobj_name}}
#define _INCLUDED_{{codeobj_name}}
void _run_{{codeobj_name}}();
{% if b200_counted %}
void _run_{{codeobj_name}}_apply();   // apply pass of a counted pathway (see synapses.cu)
{% endif %}
#endif
{% endmacro %}


{% macro before_run_cpp_file() %}
#include "code_objects/before_run_{{codeobj_name}}.h"
#include "objects.h"
#include "b200_objects.h"
#include<cmath>
#include<iostream>
#include<climits>

void _before_run_{{codeobj_name}}()
{
    using namespace brian;
    ///// CONSTANTS ///////////
    %CONSTANTS%
    ///// POINTERS ////////////
    {{host_pointers_lines|autoindent}}
    {% block before_code %}
    // EMPTY_CODE_BLOCK  -- will be overwritten in child templates
    {% endblock %}
}
{% endmacro %}

{% macro before_run_h_file() %}
#ifndef _INCLUDED_{{codeobj_name}}_before
#define _INCLUDED_{{codeobj_name}}_before
void _before_run_{{codeobj_name}}();
#endif
{% endmacro %}

{% macro after_run_cpp_file() %}
#include "code_objects/after_run_{{codeobj_name}}.h"
#include "objects.h"
#include "b200_objects.h"
#include<cmath>
#include<iostream>
#include<climits>

void _after_run_{{codeobj_name}}()
{
    using namespace brian;
    ///// CONSTANTS ///////////
    %CONSTANTS%
    ///// POINTERS ////////////
    {{host_pointers_lines|autoindent}}
    {% block after_code %}
    // EMPTY_CODE_BLOCK  -- will be overwritten in child templates
    {% endblock %}
}
{% endmacro %}

{% macro after_run_h_file() %}
#ifndef _INCLUDED_{{codeobj_name}}_after
#define _INCLUDED_{{codeobj_name}}_after
void _after_run_{{codeobj_name}}();
#endif
{% endmacro %}
