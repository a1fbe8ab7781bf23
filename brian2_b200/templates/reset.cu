{# USES_VARIABLES { N } #}
{# Resetter: brian2/devices/cpp_standalone/templates/reset.cpp:3-23.  Every CTA resets the
   neurons of ITS OWN segment of this step's spike list (the ones its thresholder just found),
   so the reset is element-private like the state update: no grid barrier is needed for it. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    const b200::EventSpaceDev& _es = _A._es{{get_array_name(eventspace_variable, access_data=False)}};
    const size_t _slot = (size_t)b200::ring_index(_clks.{{b200_clock}}.timestep, _es.slots);
    const int _segi = _ctx.rank * _ctx.gnb + _ctx.gbid;
    const int32_t _num_events = _es.cnt[_slot * (size_t)_es.nseg + _segi] & 0xffff;
    const unsigned long long* _events = _es.ids + _slot * (size_t)_es.N + _es.seg_start[_segi];
    {{scalar_code|autoindent}}
    for (int32_t _index_events = threadIdx.x; _index_events < _num_events; _index_events += b200::kBlock)
    {
        const int _idx = (int)(unsigned int)_events[_index_events];
        const int _vectorisation_idx = _idx;
        {% if b200_uses_rng %}
        b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _clks.{{b200_clock}}.timestep);
        {% endif %}
        {{vector_code|autoindent}}
    }
{% endblock %}
