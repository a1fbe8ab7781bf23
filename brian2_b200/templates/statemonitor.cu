{# USES_VARIABLES { t, _clock_t, _indices, N } #}
{# WRITES_TO_READ_ONLY_VARIABLES { t, N } #}
{# State monitor: brian2/devices/cpp_standalone/templates/statemonitor.cpp:5-42.  One row of
   the row-major (steps x n_rec) device buffer per call; the host sizes the buffer for the
   whole launch beforehand, so there is no resize in the loop.  A recorded element is read by
   the CTA that owns it in the state updater's partition (element-private access: no grid
   barrier against the state update that follows); on several GPUs a rank fills the columns of
   the neurons it owns and the host merges the columns after the run. #}
{% extends 'common_group.cu' %}
{% block maincode %}
    {% if b200_source_size is not none %}
    // Does this CTA own any recorded element?  Decided once per launch (the recorded indices do
    // not change inside a run); CTAs that own none leave at once in every later step.
    const b200::Slice _mine = b200::owned_cta((int64_t){{b200_source_size}}, _ctx);
    {% if b200_memo_slot is not none %}
    int* _memo = b200::monitor_memo() + {{b200_memo_slot}};
    if (*_memo < 0)
    {
        int _any = 0;
        for (int _i = threadIdx.x; _i < (int)_num_indices; _i += b200::kBlock)
        {
            const int _idx = {{_indices}}[_i];
            if (_idx >= _mine.lo && _idx < _mine.hi) _any = 1;
        }
        _any = __syncthreads_or(_any);
        if (threadIdx.x == 0) *_memo = _any ? 0 : 1;
        __syncthreads();
    }
    if (*_memo == 1 && _ctx.bid != 0) return;
    {% endif %}
    {% endif %}
    const int _par = (int)(_clks.{{b200_clock}}.timestep & 1);
    long long* _monN = _A._monN_{{owner.name}};
    const long long _row = _monN[_par];
    if (_ctx.bid == 0 && threadIdx.x == 0)
    {
        _A.{{b200_field(variables['t'])}}[_row] = {{_clock_t}};
        _monN[1 - _par] = _row + 1;
        {{N}} = (int32_t)(_row + 1);
    }
    // scalar code
    {{scalar_code|autoindent}}
    {% if b200_source_size is not none %}
    for (int _i = threadIdx.x; _i < (int)_num_indices; _i += b200::kBlock)
    {
        const int _idx = {{_indices}}[_i];
        if (_idx < _mine.lo || _idx >= _mine.hi) continue;
    {% else %}
    {# source is not a whole NeuronGroup (synapses, subgroup): no element ownership to exploit;
       the recorded entries are spread over the CTAs and the barrier analysis treats the reads as
       shared #}
    for (int _i = _ctx.bid * b200::kBlock + threadIdx.x; _i < (int)_num_indices; _i += _ctx.nb * b200::kBlock)
    {
        const int _idx = {{_indices}}[_i];
    {% endif %}
        const int _vectorisation_idx = _idx;
        {{vector_code|autoindent}}
        {% for varname, var in _recorded_variables | dictsort %}
        _A.{{b200_field(var)}}[_row * (long long)_num_indices + _i] = _to_record_{{varname}};
        {% endfor %}
    }
{% endblock %}
