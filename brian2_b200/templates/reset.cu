{# USES_VARIABLES { N } #}
{# Resetter: brian2/devices/cpp_standalone/templates/reset.cpp:3-23 -- scatter over this step's
   spike list (grid-stride; one spike per thread). #}
{% extends 'common_group.cu' %}
{% block maincode %}
    {% set _eventspace = get_array_name(eventspace_variable) %}
    const int32_t* _events = {{_eventspace}};
    const int32_t _num_events = {{_eventspace}}[N];
    {{scalar_code|autoindent}}
    for (int32_t _index_events = _ctx.bid * b200::kBlock + threadIdx.x; _index_events < _num_events;
         _index_events += _ctx.nb * b200::kBlock)
    {
        const int _idx = _events[_index_events];
        const int _vectorisation_idx = _idx;
        {% if b200_uses_rng %}
        b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _clks.{{b200_clock}}.timestep);
        {% endif %}
        {{vector_code|autoindent}}
    }
{% endblock %}
