{# USES_VARIABLES { _synaptic_pre, _synaptic_post, rand,
                    N_incoming, N_outgoing, N,
                    N_pre, N_post, _source_offset, _target_offset } #}
{# WRITES_TO_READ_ONLY_VARIABLES { _synaptic_pre, _synaptic_post,
                                   N_incoming, N_outgoing, N} #}
{# Synapse creation ON THE DEVICE ("sharded construction", prefs.devices.b200.construction):
   counterpart of brian2/devices/cpp_standalone/templates/synapses_create_generator.cpp.
   One thread per outer index (row) walks the row's candidates with the row's own Philox
   stream (csrc/b200_connect.cuh); a rank keeps the synapses whose postsynaptic neuron it owns.
   Same four abstract-code blocks as the reference (setup_iterator, generator_expr, create_cond,
   update; synapses.py:2138-2200), each pasted into its own scope with the values the next
   block needs handed on through row-scope variables.  Statistically equivalent to the
   reference's connect(), NOT draw-for-draw identical (different generator); independent of the
   number of GPUs by construction. #}
{% macro cpp_file() %}
// ===== code object {{codeobj_name}} (template: synapses_create_generator, on device) =====
namespace _co_{{codeobj_name}} {
    ////// SUPPORT CODE ///////
    {{support_code_lines|autoindent}}
    struct Scal {
        {% for block in ['setup_iterator', 'generator_expr', 'create_cond', 'update'] %}
        {% for ctype, name in b200_scalar_members[block] %}
        {{ctype}} {{name}};
        {% endfor %}
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
    {% for block in ['setup_iterator', 'generator_expr', 'create_cond', 'update'] %}
    {{b200_scalar_host[block]|autoindent}}
    {% endfor %}
    _sc._unused = 0;
}

__global__ void __launch_bounds__(256)
_kernel_{{codeobj_name}}(const b200::ConnectArgs _args, const _co_{{codeobj_name}}::Scal _sc)
{
    using namespace _co_{{codeobj_name}};
    ///// CONSTANTS ///////////
    %CONSTANTS_DEV%
    const int64_t _N_pre = {{constant_or_scalar('N_pre', variables['N_pre'])}};
    const int64_t _N_post = {{constant_or_scalar('N_post', variables['N_post'])}};
    {% for pointer, start in b200_identity_arrays %}
    const b200::IdentityIndex {{pointer}}{ {{start}} };   // arange array: the value is the index
    {% endfor %}
    // scalar code
    {{scalar_code['setup_iterator']|autoindent}}
    {{scalar_code['generator_expr']|autoindent}}
    {{scalar_code['create_cond']|autoindent}}
    {{scalar_code['update']|autoindent}}
    for (long long _row = (long long)blockIdx.x * blockDim.x + threadIdx.x; _row < _args.n_outer;
         _row += (long long)gridDim.x * blockDim.x)
    {
        const int32_t _{{outer_index}} = (int32_t)_row;
        const int _vectorisation_idx = (int)_row;
        b200::Rng _rng = b200::rng_init(_args.seed, _args.stream, _row, 0);
        b200::RowSink _sink(_args, _row);
        int32_t _raw_pre_idx = 0, _raw_post_idx = 0;
        _raw{{outer_index_array}} = _{{outer_index}} + (int32_t)({{outer_index_offset}});
        bool _row_cond = true;
        {% if not result_index_condition %}
        {   // condition that only depends on the outer index: decided once per row
            {{vector_code['create_cond']|autoindent}}
            _row_cond = _cond;
        }
        {% endif %}
        b200::CandidateIter _it;
        _it.done = true;
        if (_row_cond)
        {
            {{vector_code['setup_iterator']|autoindent}}
            {% if iterator_func == 'range' %}
            _it.init_range((long)_iter_low, (long)_iter_high, (long)_iter_step);
            {% else %}
            _it.init_sample((long)_iter_low, (long)_iter_high, (long)_iter_step, (double)_iter_p);
            {% endif %}
        }
        long _cand;
        while (_it.next(_cand, _rng))
        {
            const int32_t {{inner_variable}} = (int32_t)_cand;
            long _b200_result;
            int32_t _b200_outer_arr;
            {
                {{vector_code['generator_expr']|autoindent}}
                _b200_result = (long)_{{result_index}};
                _b200_outer_arr = {{outer_index_array}};
            }
            const int32_t _{{result_index}} = (int32_t)_b200_result;
            const int32_t {{outer_index_array}} = _b200_outer_arr;
            const bool _b200_in_range = _b200_result >= 0 && _b200_result < (long)_{{result_index_size}};
            _raw{{result_index_array}} = _{{result_index}} + (int32_t)({{result_index_offset}});
            bool _b200_keep = true;
            {% if result_index_condition %}
            {% if result_index_used %}
            if (!_b200_in_range)
            {
                {% if skip_if_invalid %}
                continue;
                {% else %}
                *_args.error = 1;
                break;
                {% endif %}
            }
            {% endif %}
            {
                {{vector_code['create_cond']|autoindent}}
                _b200_keep = _cond;
            }
            {% endif %}
            if (!_b200_keep) continue;
            if (!_b200_in_range)
            {
                {% if skip_if_invalid %}
                continue;
                {% else %}
                *_args.error = 1;
                break;
                {% endif %}
            }
            {
                {{vector_code['update']|autoindent}}
                _sink.emit((int32_t)_pre_idx, (int32_t)_post_idx, (int)_n);
            }
        }
        _sink.finish(_row);
    }
}

void _run_{{codeobj_name}}()
{
    using namespace brian;
    _co_{{codeobj_name}}::Scal _sc;
    _hostscal_{{codeobj_name}}(_sc);
    ///// CONSTANTS ///////////
    %CONSTANTS%
    const int64_t _N_pre = {{b200_host_constant_or_scalar('N_pre', variables['N_pre'])}};
    const int64_t _N_post = {{b200_host_constant_or_scalar('N_post', variables['N_post'])}};
    const long long _n_outer = (long long)_{{outer_index_size}};
    // every connect() call of a script gets its own family of streams
    static unsigned int _b200_calls = 0;
    const unsigned int _stream = {{b200_stream_id}}u + 0x9E3779B9u * (_b200_calls++);
    const b200::ConnectResult _res = b200::connect_on_device(
        [&](const b200::ConnectArgs& _a) {
            const int _grid = (int)std::min<long long>(std::max<long long>((_n_outer + 255) / 256, 1), 148 * 64);
            _kernel_{{codeobj_name}}<<<_grid, 256, 0, b200::state().stream>>>(_a, _sc);
        },
        _stream, _n_outer, (int64_t){{b200_post_parent_size}},
        (size_t)(_N_pre + _source_offset), (size_t)(_N_post + _target_offset),
        {{_dynamic__synaptic_pre}}, {{_dynamic__synaptic_post}},
        {{_dynamic_N_incoming}}, {{_dynamic_N_outgoing}});
    (void)_res;
    // every per-synapse variable follows the new number of (local) synapses
    const int32_t _newsize = (int32_t){{_dynamic__synaptic_pre}}.size();
    {% for varname in owner._registered_variables | variables_to_array_names(access_data=False) | sort %}
    {{varname}}.resize(_newsize);
    {% endfor %}
    {{get_array_name(variables['N'], access_data=False)}}[0] = _newsize;
}
{% endmacro %}


{% macro h_file() %}
#ifndef _INCLUDED_{{codeobj_name}}
#define _INCLUDED_{{codeobj_name}}
void _run_{{codeobj_name}}();
#endif
{% endmacro %}
