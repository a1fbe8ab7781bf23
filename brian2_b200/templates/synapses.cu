{# USES_VARIABLES { _synaptic_pre } #}
{# Synaptic effect (on_pre / on_post): brian2/devices/cpp_standalone/templates/synapses.cpp:11-50.
   Instead of peeking a bucket of synapse ids, every delay bin d reads the spike list emitted d
   steps ago from the source's spike ring and walks the CSR rows of those neurons: one warp per
   (spike, delay bin) row, lanes stride through the row (coalesced index / weight reads,
   atomics on the postsynaptic side).  When the abstract code is order dependent the walk is
   done by one thread in the reference's delivery order (largest delay first, spiking neuron
   ascending, synapse index ascending; spikequeue.h:157-190). #}
{% extends 'common_group.cu' %}
{% block maincode %}
    const b200::PathwayDev& _pw = _A._pw_{{pathway.name}};
    const int64_t _b200_timestep = _clks.{{b200_clock}}.timestep;
    // scalar code
    {{scalar_code|autoindent}}
    {% if b200_serial %}
    if (_ctx.bid == 0 && threadIdx.x == 0)
    {
        for (int _bin = _pw.nbins - 1; _bin >= 0; --_bin)
        {
            const int32_t* _spk = b200::ring_slot(_pw.ring, _pw.ring_slots, _pw.ring_stride,
                                                  _b200_timestep - _pw.bin_delay[_bin]);
            const int _nspk = _spk[_pw.ring_stride - 1];
            const int* _rp = _pw.rowptr + (size_t)_bin * (_pw.nsrc + 1);
            for (int _s = 0; _s < _nspk; ++_s)
            {
                const int _src = _spk[_s] - _pw.src_start;
                if (_src < 0 || _src >= _pw.nsrc) continue;
                const int _beg = _rp[_src], _end = _rp[_src + 1];
                *_pw.events += (unsigned long long)(_end - _beg);
                for (int _k = _beg; _k < _end; ++_k)
                {
                    const int _idx = _pw.identity ? _k : _pw.syn_ids[_k];
                    const int _vectorisation_idx = _idx;
                    {% if b200_uses_rng %}
                    b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _b200_timestep);
                    {% endif %}
                    {{vector_code|autoindent}}
                }
            }
        }
    }
    {% else %}
    const int _lane = threadIdx.x & 31;
    const int _gwarp = _ctx.bid * b200::kWarps + (threadIdx.x >> 5);
    const int _nwarps = _ctx.nb * b200::kWarps;
    for (int _bin = 0; _bin < _pw.nbins; ++_bin)
    {
        const int32_t* _spk = b200::ring_slot(_pw.ring, _pw.ring_slots, _pw.ring_stride,
                                              _b200_timestep - _pw.bin_delay[_bin]);
        const int _nspk = _spk[_pw.ring_stride - 1];
        const int* _rp = _pw.rowptr + (size_t)_bin * (_pw.nsrc + 1);
        for (int _s = _gwarp; _s < _nspk; _s += _nwarps)
        {
            const int _src = _spk[_s] - _pw.src_start;
            if (_src < 0 || _src >= _pw.nsrc) continue;
            const int _beg = _rp[_src], _end = _rp[_src + 1];
            if (_lane == 0 && _end > _beg)
                atomicAdd(_pw.events, (unsigned long long)(_end - _beg));
            for (int _k = _beg + _lane; _k < _end; _k += 32)
            {
                const int _idx = _pw.identity ? _k : _pw.syn_ids[_k];
                const int _vectorisation_idx = _idx;
                {% if b200_uses_rng %}
                b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _b200_timestep);
                {% endif %}
                {{vector_code|autoindent}}
            }
        }
    }
    {% endif %}
{% endblock %}
