{# USES_VARIABLES { _spikespace, neuron_index, _timebins, _period_bins, _lastindex, t_in_timesteps, N } #}
{# WRITES_TO_READ_ONLY_VARIABLES { _lastindex } #}
{# SpikeGeneratorGroup: brian2/devices/cpp_standalone/templates/spikegenerator.cpp:3-32.  The
   reference walks the (time-sorted) spike table from a running cursor `_lastindex`.  Here the
   spikes of the current time bin are found by binary search (no sequential state), every CTA
   marks the ones that belong to the neurons IT owns in a shared-memory bit set, and the bits
   are published exactly like a thresholder's result (ascending ids in the CTA's segment of the
   spike ring).  `_lastindex` is still maintained for the host (before_run reads it). #}
{% extends 'common_group.cu' %}
{% block maincode %}
    const int64_t _N = N;
    const b200::EventSpaceDev& _es = _A._es{{get_array_name(variables['_spikespace'], access_data=False)}};
    const int32_t _the_period = {{_period_bins}};
    int32_t _timebin = (int32_t){{t_in_timesteps}};
    if (_the_period > 0) _timebin %= _the_period;
    // [_first, _last): entries of the table that fall into this time bin
    const int _ntab = (int)_num_timebins;
    const int _first = b200::lower_bound_i32({{_timebins}}, _ntab, _timebin);
    const int _last = b200::lower_bound_i32({{_timebins}}, _ntab, _timebin + 1);
    __shared__ unsigned int _s_bits[b200::kMaxOwnedIters * b200::kWarps];   // 1 bit per owned element
    const b200::Slice _cta = b200::owned_cta(_N, _ctx);
    const b200::Slice _sl = b200::owned_slice(_N, _ctx);
    for (int _w = threadIdx.x; _w < b200::kMaxOwnedIters * b200::kWarps; _w += b200::kBlock) _s_bits[_w] = 0u;
    __syncthreads();
    for (int _k = _first + threadIdx.x; _k < _last; _k += b200::kBlock)
    {
        const int64_t _id = {{neuron_index}}[_k];
        if (_id >= _cta.lo && _id < _cta.hi)
        {
            const int _rel = (int)(_id - _cta.lo);
            atomicOr(&_s_bits[_rel >> 5], 1u << (_rel & 31));
        }
    }
    __syncthreads();
    const int _niter = (int)((_sl.hi - _sl.lo + 31) >> 5);
    unsigned long long _mask = 0ULL;
    {
        int _k = 0;
        for (int64_t _i64 = _sl.lo + (threadIdx.x & 31); _i64 < _sl.hi; _i64 += 32, ++_k)
        {
            const int _rel = (int)(_i64 - _cta.lo);
            if ((_s_bits[_rel >> 5] >> (_rel & 31)) & 1u) _mask |= (1ULL << _k);
        }
    }
    b200::publish_owned(_mask, _niter, _ctx, _es, _clks.{{b200_clock}}.timestep);
    if (_ctx.gbid == 0 && threadIdx.x == 0) {{_lastindex}} = _last;
{% endblock %}

{% block after_code %}
    {{_spikespace}}[N] = 0;   // host mirror, as the thresholder's after_run block (threshold.cpp:34-37)
{% endblock %}
