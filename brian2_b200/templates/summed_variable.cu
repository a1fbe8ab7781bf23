{# USES_VARIABLES { N } #}
{# Summed variables (`I_post = g*(v_pre - v_post) : amp (summed)`):
   brian2/devices/cpp_standalone/templates/summed_variable.cpp:4-29.  The reference zeroes the
   target and then adds every synapse's value in synapse-index order.  Here the sum is a GATHER:
   a CSR by target element (built once in before_run, synapse indices ascending per row) lets
   one thread own one target element and add its synapses' values in exactly the reference's
   order -- no zeroing pass, no atomics, bit-identical sums. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    {% set _target_var_array = get_array_name(_target_var) %}
    const b200::TargetIndexDev& _ti = _A._sv_{{owner.name}}{{get_array_name(_target_var, access_data=False)}};
    // scalar code
    {{scalar_code|autoindent}}
    {% if b200_target_whole_group %}
    B200_FOR_OWNED(_t64, _ti.n_targets, _ctx)
    {
        const int _target_idx = (int)_t64;
    {% else %}
    for (int _target_idx = _ctx.bid * b200::kBlock + threadIdx.x; _target_idx < _ti.n_targets;
         _target_idx += _ctx.nb * b200::kBlock)
    {
    {% endif %}
        const int _beg = _ti.rowptr[_target_idx], _end = _ti.rowptr[_target_idx + 1];
        {{c_data_type(_target_var.dtype)}} _sum = 0;
        for (int _k = _beg; _k < _end; ++_k)
        {
            const int _idx = _ti.syn_ids[_k];
            const int _vectorisation_idx = _idx;
            {% if b200_uses_rng %}
            b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _clks.{{b200_clock}}.timestep);
            {% endif %}
            {{vector_code|autoindent}}
            _sum += _synaptic_var;
        }
        {{_target_var_array}}[_target_idx + {{_target_start}}] = _sum;
    }
{% endblock %}

{% block before_code %}
    {% set _index_array = get_array_name(_index_var, access_data=False) %}
    {
        std::vector<int32_t>& _b200_index = {{_index_array}};
        const int _target_size = {{b200_host_constant_or_scalar(_target_size_name, variables[_target_size_name])}};
        _b200_sv_{{owner.name}}{{get_array_name(_target_var, access_data=False)}}.prepare(_b200_index.empty() ? 0 : &_b200_index[0], _b200_index.size(),
                                       {{_target_start}}, _target_size);
    }
{% endblock %}
