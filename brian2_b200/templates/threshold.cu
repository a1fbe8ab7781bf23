{# USES_VARIABLES { N } #}
{# Thresholder: brian2/devices/cpp_standalone/templates/threshold.cpp:3-37.  The reference's loop
   is serial (`_count++`); here the condition is evaluated with the same element->lane mapping
   as the state updater and every CTA writes the ids of its own neurons, ascending (warp
   ballots + one intra-CTA scan), into its own segment of the current slot of the event space:
   no communication between CTAs, and "pushing" the spikes to the synaptic pathways costs
   nothing.  On several GPUs the same stores also go to the peers' rings over NVLink. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    {% set _eventspace = get_array_name(eventspace_variable) %}
    {{scalar_code|autoindent}}
    const int64_t _N = N;
    const b200::Slice _sl = b200::owned_slice(_N, _ctx);
    const int _niter = (int)((_sl.hi - _sl.lo + 31) >> 5);
    unsigned long long _mask = 0ULL;
    {
        int _k = 0;
        for (int64_t _i64 = _sl.lo + (threadIdx.x & 31); _i64 < _sl.hi; _i64 += 32, ++_k)
        {
            {
                const int _idx = (int)_i64;
                const int _vectorisation_idx = _idx;
                {% if b200_uses_rng %}
                b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _clks.{{b200_clock}}.timestep);
                {% endif %}
                {{vector_code|autoindent}}
                if (_cond)
                {
                    _mask |= (1ULL << _k);
                    {% if _uses_refractory %}
                    {{not_refractory}}[_idx] = false;
                    {{lastspike}}[_idx] = {{t}};
                    {% endif %}
                }
            }
        }
    }
    b200::publish_owned(_mask, _niter, _ctx, _A._es{{get_array_name(eventspace_variable, access_data=False)}},
                        _clks.{{b200_clock}}.timestep);
{% endblock %}

{% block after_code %}
    {% set _eventspace = get_array_name(eventspace_variable) %}
    {{_eventspace}}[N] = 0;  // host mirror; the array has N+1 elements (threshold.cpp:34-37)
{% endblock %}
