{# USES_VARIABLES { _synaptic_pre } #}
{# Synaptic effect (on_pre / on_post): brian2/devices/cpp_standalone/templates/synapses.cpp:11-50.
   Instead of peeking a bucket of synapse ids, every delay bin d reads the spike list emitted d
   steps ago from the source's event space and walks the CSR rows of those neurons: one warp per
   (spike, delay bin) row, lanes stride through the row (coalesced index / weight reads,
   atomics on the postsynaptic side).  The list of the current step (delay 0) is read straight
   from the thresholder's per-CTA segments (b200::view_*), older steps from their compacted
   form.  When the abstract code is order dependent the walk is done by one thread in the
   reference's delivery order (largest delay first, spiking neuron ascending, synapse index
   ascending; spikequeue.h:157-190). #}
{% extends 'common_group.cu' %}
{% block maincode %}
    {% set _es = '_A._es' + get_array_name(pathway.source.variables[pathway.eventspace_name], access_data=False) %}
    const b200::PathwayDev& _pw = _A._pw_{{pathway.name}};
    const b200::EventSpaceDev& _es = {{_es}};
    const int64_t _b200_timestep = _clks.{{b200_clock}}.timestep;
    // scalar code
    {{scalar_code|autoindent}}
    {% if b200_serial %}
    {# serial fallback: needs the current step compacted, see B200Device._plan_barriers #}
    if (_ctx.bid == 0 && threadIdx.x == 0)
    {
        for (int _bin = _pw.nbins - 1; _bin >= 0; --_bin)
        {
            const int32_t* _spk = b200::compact_slot(_es, _b200_timestep - _pw.bin_delay[_bin]);
            const int _nspk = _spk[_es.N];
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
    unsigned long long _nev = 0ULL;
    for (int _bin = 0; _bin < _pw.nbins; ++_bin)
    {
        const int _delay = _pw.bin_delay[_bin];
        const int* _rp = _pw.rowptr + (size_t)_bin * (_pw.nsrc + 1);
        b200::SpikeView _view;
        const int32_t* _spk = 0;
        int _nspk;
        if (_delay == 0)
        {
            _view = b200::view_build(_es, _b200_timestep, _ctx, false, _A._ctrl);
            _nspk = _view.total;
        }
        else
        {
            _spk = b200::compact_slot(_es, _b200_timestep - _delay);
            _nspk = _spk[_es.N];
        }
        for (int _s = _gwarp; _s < _nspk; _s += _nwarps)
        {
            const int _src = (_delay == 0 ? b200::view_id(_view, _s) : _spk[_s]) - _pw.src_start;
            if (_src < 0 || _src >= _pw.nsrc) continue;
            const int _beg = _rp[_src], _end = _rp[_src + 1];
            _nev += (unsigned long long)(_end - _beg);
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
    // delivered synaptic events (the benchmark metric): one atomic per warp per step
    if (_lane == 0 && _nev) atomicAdd(_pw.events, _nev);
    {% endif %}
{% endblock %}
