{# USES_VARIABLES { N, _clock_t, count, _source_start, _source_stop} #}
{# WRITES_TO_READ_ONLY_VARIABLES { N, count } #}
{# Spike/event monitor: brian2/devices/cpp_standalone/templates/spikemonitor.cpp:6-51.
   The recorded ids are a contiguous run of the (ascending) spike list, so every spike's output
   position is known without atomics: N_old + (g - first).  The list is read straight from the
   thresholder's segments (b200::view_*); a binary search is only needed when the monitor
   watches a subgroup.  The running total N is double buffered by step parity so that all CTAs
   can read it while one thread publishes the new value.  Storage is a device append buffer;
   the host grows it between launches when the kernel reports that fewer than one step's worst
   case of free slots are left.  On several GPUs every rank records the spikes of its own
   neurons; the host merges the per-rank records by (t, i) after the run. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    const b200::EventSpaceDev& _es = _A._es{{get_array_name(eventspace_variable, access_data=False)}};
    // (issued first: this load is in flight while the view is built)
    const int _par = (int)(_clks.{{b200_clock}}.timestep & 1);
    long long* _monN = _A._monN_{{owner.name}};
    const long long _N_old = b200::ld_relaxed_s64(_monN + _par);
    const b200::SpikeView _view = b200::view_build(_es, _clks.{{b200_clock}}.timestep, _ctx, true, _A._ctrl);
    int _start_idx = 0, _end_idx = _view.total;
    if ((int)_source_start > 0 || (int)_source_stop < _es.N)
    {
        _start_idx = b200::view_count_below(_view, _es, (int)_source_start);
        _end_idx = b200::view_count_below(_view, _es, (int)_source_stop);
    }
    const int _num_events = _end_idx - _start_idx;
    if (_num_events > 0)
    {
        const int _vectorisation_idx = 1;
        {{scalar_code|autoindent}}
        for (int _j = _start_idx + _ctx.bid * b200::kBlock + threadIdx.x; _j < _end_idx;
             _j += _ctx.nb * b200::kBlock)
        {
            const int _idx = b200::view_id(_view, _j);
            const int _vectorisation_idx = _idx;
            {{vector_code|autoindent}}
            const long long _out = _N_old + (_j - _start_idx);
            {% for varname, var in record_variables | dictsort %}
            _A.{{b200_field(var)}}[_out] = _to_record_{{varname}};
            {% endfor %}
            // (the ids of one step are distinct: a reduction without return value, no round trip)
            atomicAdd(&{{count}}[_idx - _source_start], 1);
        }
    }
    if (_ctx.bid == 0 && threadIdx.x == 0)
    {
        const long long _N_new = _N_old + _num_events;
        _monN[1 - _par] = _N_new;
        {{N}} = (int32_t)_N_new;
        {% if record_variables %}
        {% set _first = (record_variables | dictsort | first)[1] %}
        // (a stop raised here takes effect at the next grid barrier, which may be one step away)
        if (_N_new + 2LL * (long long)(_source_stop - _source_start) > (long long)_A._cap{{b200_field(_first)}})
        {
            _A._ctrl->overflow = 1;
            b200::raise_stop(_A._ctrl);
        }
        {% endif %}
    }
{% endblock %}
