{# USES_VARIABLES { N } #}
{# ALLOWS_SCALAR_WRITE #}
{# Per-element state update: brian2/devices/cpp_standalone/templates/stateupdate.cpp:5-22.
   One element per lane, warp-contiguous owned slices => 256 B coalesced fp64 requests. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    // scalar code (hoisted `_lio_*` come from _sc)
    {{scalar_code|autoindent}}
    const int64_t _N = {{constant_or_scalar('N', variables['N'])}};
    B200_FOR_OWNED(_i64, _N, _ctx)
    {
        const int _idx = (int)_i64;
        const int _vectorisation_idx = _idx;
        {% if b200_uses_rng %}
        b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _clks.{{b200_clock}}.timestep);
        {% endif %}
        {{vector_code|autoindent}}
    }
{% endblock %}
