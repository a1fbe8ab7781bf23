{# USES_VARIABLES { N, rate, t, _spikespace, _clock_t, _clock_dt,
                    _num_source_neurons, _source_start, _source_stop } #}
{# WRITES_TO_READ_ONLY_VARIABLES { N } #}
{# Population rate monitor: brian2/devices/cpp_standalone/templates/ratemonitor.cpp:6-36.
   CTA 0 sums the segment counts of this step (b200::view_build).  On several GPUs every rank
   stores the NUMBER of spikes of its own neurons in `rate`; the host adds the ranks up and
   applies the reference's `1.0*n/dt/N` once (exactly the reference's arithmetic). #}
{% extends 'common_group.cu' %}
{% block maincode %}
    if (_ctx.bid == 0)
    {
        const b200::EventSpaceDev& _es = _A._es{{get_array_name(variables['_spikespace'], access_data=False)}};
        const b200::SpikeView _view = b200::view_build(_es, _clks.{{b200_clock}}.timestep, _ctx, true, _A._ctrl);
        if (threadIdx.x == 0)
        {
            int _num_spikes = _view.total;
            if ((int)_source_start > 0 || (int)_source_stop < _es.N)
                _num_spikes = b200::view_count_below(_view, _es, (int)_source_stop)
                              - b200::view_count_below(_view, _es, (int)_source_start);
            const int _par = (int)(_clks.{{b200_clock}}.timestep & 1);
            long long* _monN = _A._monN_{{owner.name}};
            const long long _n = _monN[_par];
            _monN[1 - _par] = _n + 1;
            if (_ctx.world > 1)
                _A.{{b200_field(variables['rate'])}}[_n] = (double)_num_spikes;
            else
                _A.{{b200_field(variables['rate'])}}[_n] = 1.0*_num_spikes/{{_clock_dt}}/_num_source_neurons;
            _A.{{b200_field(variables['t'])}}[_n] = {{_clock_t}};
            {{N}} = (int32_t)(_n + 1);
        }
    }
{% endblock %}
