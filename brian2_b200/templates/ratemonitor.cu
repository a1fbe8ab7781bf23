{# USES_VARIABLES { N, rate, t, _spikespace, _clock_t, _clock_dt,
                    _num_source_neurons, _source_start, _source_stop } #}
{# WRITES_TO_READ_ONLY_VARIABLES { N } #}
{# Population rate monitor: brian2/devices/cpp_standalone/templates/ratemonitor.cpp:6-36.
   One thread: two binary searches on the ascending spike list replace the linear scans. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    if (_ctx.bid == 0 && threadIdx.x == 0)
    {
        const int32_t* _events = {{_spikespace}};
        const int _num_all = _events[_num_spikespace - 1];
        const int _start_idx = b200::lower_bound_i32(_events, _num_all, (int)_source_start);
        const int _end_idx = b200::lower_bound_i32(_events, _num_all, (int)_source_stop);
        const int _num_spikes = _end_idx - _start_idx;
        const int _par = (int)(_clks.{{b200_clock}}.timestep & 1);
        long long* _monN = _A._monN_{{owner.name}};
        const long long _n = _monN[_par];
        _monN[1 - _par] = _n + 1;
        _A.{{b200_field(variables['rate'])}}[_n] = 1.0*_num_spikes/{{_clock_dt}}/_num_source_neurons;
        _A.{{b200_field(variables['t'])}}[_n] = {{_clock_t}};
        {{N}} = (int32_t)(_n + 1);
    }
{% endblock %}
