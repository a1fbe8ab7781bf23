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
                    const int _b200_src_idx = _src + _pw.src_start, _b200_tgt_idx = _pw.csr_target[_k];
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
    {% if b200_counted %}
    {% if not b200_counted.dual %}
    // counted pathway: the delivery only counts the events per target (integer reductions);
    // the owner of every target applies them afterwards (_dev_{{codeobj_name}}_apply below)
    int* _b200_hits = _pw.hits + (size_t)(_b200_timestep % _pw.hits_slots) * (size_t)_pw.hits_n;
    if (_pw.forward)
    {
        // ---- FORWARD delivery (every delay >= 1 step; CSR by (source, delay bin), see
        // Pathway::prepare): the spikes of the PREVIOUS step are delivered now, each row once and
        // contiguous, into the counters of the steps in which the synapses are due.
        const int _lane = threadIdx.x & 31;
        const int _gwarp = _ctx.bid * b200::kWarps + (threadIdx.x >> 5);
        const int _nwarps = _ctx.nb * b200::kWarps;
        const int _R = _pw.hits_slots, _nb = _pw.nbins;
        const int _bdelay = _lane < _nb ? __ldg(_pw.bin_delay + _lane) : 0;
        const b200::SpikeView _view = b200::view_build(_es, _b200_timestep - 1, _ctx, false, _A._ctrl);
        const int _nrows = _view.total;
        const int _base_slot = (int)((_b200_timestep - 1) % _R);      // slot of delay 0 (never used: delays >= 1)
        unsigned long long _nev = 0ULL;
        // fewer rows than warps: _k warps share a row; else a warp takes whole rows
        const int _k = _nrows > 0 && _nrows < _nwarps ? _nwarps / _nrows : 1;
        for (int _r = _gwarp / _k; _r < _nrows; _r += max(1, _nwarps / _k))
        {
            const int _part = _gwarp - (_gwarp / _k) * _k;
            const int _src = b200::view_id(_view, _r) - _pw.src_start;
            if (_src < 0 || _src >= _pw.nsrc) continue;
            const int* _rp = _pw.rowptr + (size_t)_src * (size_t)(_nb + 1);
            const int _beg = __ldg(_rp), _end = __ldg(_rp + _nb);
            if (_part == 0) _nev += (unsigned long long)(_end - _beg);
            const int _per = (((_end - _beg + _k - 1) / _k) + 31) & ~31;
            const int _a = _beg + _part * _per, _b = min(_end, _a + _per);
            for (int _kb = _a + _lane; _kb - _lane < _b; _kb += 32 * 4)
            {
                unsigned int _w[4];
                #pragma unroll
                for (int _u = 0; _u < 4; ++_u)
                    _w[_u] = _kb + 32 * _u < _b ? (unsigned int)b200::ld_index(_pw.csr_target + _kb + 32 * _u) : 0xffffffffu;
                #pragma unroll
                for (int _u = 0; _u < 4; ++_u)
                {
                    const unsigned int _bin = _w[_u] == 0xffffffffu ? 0u : (_w[_u] >> 27);
                    int _slot = _base_slot + __shfl_sync(0xffffffffu, _bdelay, (int)_bin);
                    if (_slot >= _R) _slot -= _R;
                    if (_w[_u] != 0xffffffffu)
                        atomicAdd(_pw.hits + (size_t)_slot * (size_t)_pw.hits_n + (_w[_u] & 0x7ffffffu), 1);
                }
            }
        }
        if (_lane == 0 && _nev) atomicAdd(_pw.events, _nev);
        return;
    }
    {% endif %}
    if (_pw.tileptr) return;    // dense rows: counted AND applied by the owners of the targets (apply pass)
    {% endif %}
    const int _lane = threadIdx.x & 31;
    const int _gwarp = _ctx.bid * b200::kWarps + (threadIdx.x >> 5);
    const int _nwarps = _ctx.nb * b200::kWarps;
    unsigned long long _nev = 0ULL;        // events counted by the whole warp (same value in all lanes)
    unsigned long long _nev_lane = 0ULL;   // events counted lane by lane (gather mode)
    // work counter of the NEXT step (see "heavy steps" below): zeroed by one thread every step;
    // the end-of-step barrier orders it before its first use
    if (_ctx.bid == 0 && threadIdx.x == 0) _pw.tickets[(_b200_timestep + 1) & 1] = 0u;
    // parameters of the first 32 delay bins, one per lane (in flight while the view is built)
    int _bdelay0 = 0, _blr0 = 1;
    if (_lane < _pw.nbins)
    {
        _bdelay0 = __ldg(_pw.bin_delay + _lane);
        _blr0 = (__ldg(_pw.bin_maxlen + _lane) + 62) >> 5;
    }
    // The youngest list (delay 0; delay 1 when every pathway of the project is delayed) is read
    // straight from the thresholder's segments
    const int _segd = _pw.seg_delay;
    b200::SpikeView _view;
    _view.total = 0;
    if (_segd >= 0)
        _view = b200::view_build(_es, _b200_timestep - _segd, _ctx, false, _A._ctrl);
    // Work = 128-byte lines of the packed index stream.  Every (delay bin, spike of that bin's
    // step) row is padded to the bin's longest row (`_lr` lines of 32 slots, on 32-slot boundaries:
    // a warp load is one aligned line) and the lines of all rows of all bins form one virtual
    // sequence that is dealt out to the warps of the grid in equal contiguous shares: balanced
    // to within one line whether few neurons with long rows or many with short rows fired, and
    // neighbouring warps read neighbouring lines of the same row.  Lane l of every warp holds the
    // parameters of bin _g0 + l; a line finds its bin with one ballot (the bins are independent:
    // no latency chain per bin).
    for (int _g0 = 0; _g0 < _pw.nbins; _g0 += 32)
    {
        const int _mybin = _g0 + _lane;
        int _bdelay = 0, _bn = 0, _blr = 1;
        const int32_t* _bspk = 0;
        if (_mybin < _pw.nbins)
        {
            _bdelay = _g0 == 0 ? _bdelay0 : __ldg(_pw.bin_delay + _mybin);
            _blr = _g0 == 0 ? _blr0 : ((__ldg(_pw.bin_maxlen + _mybin) + 62) >> 5);
            if (_blr < 1) _blr = 1;
            if (_bdelay == _segd)
                _bn = _view.total;
            else
            {
                _bspk = b200::compact_slot(_es, _b200_timestep - _bdelay);
                _bn = _bspk[_es.N];
            }
        }
        long long _blines = (long long)_bn * _blr;
        long long _bincl = _blines;
        #pragma unroll
        for (int _o = 1; _o < 32; _o <<= 1)
        {
            const long long _t = __shfl_up_sync(0xffffffffu, _bincl, _o);
            if (_lane >= _o) _bincl += _t;
        }
        const long long _bexcl = _bincl - _blines;
        const long long _nlines = __shfl_sync(0xffffffffu, _bincl, 31);
        if (_nlines <= 0) continue;
        int _rincl = _bn;                       // rows (spikes) of the bins, same scan
        #pragma unroll
        for (int _o = 1; _o < 32; _o <<= 1)
        {
            const int _t = __shfl_up_sync(0xffffffffu, _rincl, _o);
            if (_lane >= _o) _rincl += _t;
        }
        const int _rexcl = _rincl - _bn;
        const int _nrows = __shfl_sync(0xffffffffu, _rincl, 31);
        // (32-bit divisions whenever the numbers allow: ~10x cheaper)
        const bool _small = _nlines * (long long)_nwarps < 0xffffffffLL;
        // Heavy steps (>= 64 lines per warp, one bin group): the lines are handed out in tickets
        // of 32 lines from a device counter instead of fixed shares, so that CTAs that see a
        // slower memory system (far L2 partition) simply take fewer tickets.  The counter of the
        // NEXT step is zeroed above.
        // (only for long rows: many short rows are served by the gather mode below whatever
        // their number -- a ticket of padded lines would walk its rows one dependent chain after
        // the other)
        int _maxlr = _mybin < _pw.nbins ? _blr : 0;
        #pragma unroll
        for (int _o = 16; _o > 0; _o >>= 1) _maxlr = max(_maxlr, __shfl_xor_sync(0xffffffffu, _maxlr, _o));
        const bool _dyn = _pw.nbins <= 32 && _nlines >= 64LL * _nwarps && _maxlr > 4;
        unsigned int* _tickets = _pw.tickets + (_b200_timestep & 1);
        // Many short rows (more rows than warps, every row <= 4 lines: sharded Brunel, COBAHH at
        // high rates): GATHER mode.  A warp takes an equal share of the ROWS, 32 at a time: lane l
        // fetches the spike id and the row pointers of row l (32 chains of dependent loads in
        // flight together instead of one after the other), a warp scan of the row lengths turns
        // the 32 rows into one dense run of slots, and every lane then delivers one slot per
        // pass (all lanes busy however short the rows are; a slot finds its row with a 5-step
        // search over the scanned lengths by shuffles).
        if (!_dyn && _nrows > _nwarps && _maxlr <= 4)
        {
            const int _ra = (int)(((long long)_gwarp * _nrows) / _nwarps);
            const int _rb = (int)(((long long)(_gwarp + 1) * _nrows) / _nwarps);
            for (int _r0 = _ra; _r0 < _rb; _r0 += 32)
            {
                const bool _rv = _r0 + _lane < _rb;
                const int _r = _rv ? _r0 + _lane : _ra;
                int _blo = 0, _bhi = 32;             // bin of row _r: last bin with _rexcl <= _r
                #pragma unroll
                for (int _i = 0; _i < 5; ++_i)
                {
                    const int _mid = (_blo + _bhi) >> 1;
                    if (__shfl_sync(0xffffffffu, _rexcl, _mid) <= _r) _blo = _mid; else _bhi = _mid;
                }
                const int _s = _r - __shfl_sync(0xffffffffu, _rexcl, _blo);
                const int _delay = __shfl_sync(0xffffffffu, _bdelay, _blo);
                const int32_t* _spk = (const int32_t*)__shfl_sync(0xffffffffu, (unsigned long long)_bspk, _blo);
                int _len = 0, _rbeg = 0, _srcabs = 0;
                if (_rv)
                {
                    const int _src = (_delay == _segd ? b200::view_id(_view, _s) : _spk[_s]) - _pw.src_start;
                    if (_src >= 0 && _src < _pw.nsrc)
                    {
                        const int* _rp = _pw.rowptr + (size_t)(_g0 + _blo) * (_pw.nsrc + 1);
                        _rbeg = _rp[_src];
                        _len = _rp[_src + 1] - _rbeg;
                        _srcabs = _src + _pw.src_start;
                    }
                }
                _nev_lane += (unsigned long long)_len;
                int _lincl = _len;
                #pragma unroll
                for (int _o = 1; _o < 32; _o <<= 1)
                {
                    const int _t = __shfl_up_sync(0xffffffffu, _lincl, _o);
                    if (_lane >= _o) _lincl += _t;
                }
                const int _lexcl = _lincl - _len;
                const int _nslots = __shfl_sync(0xffffffffu, _lincl, 31);
                for (int _base = 0; _base < _nslots; _base += 32 * {{b200_gather_unroll}})
                {
                    int _b200_tg[{{b200_gather_unroll}}], _b200_sy[{{b200_gather_unroll}}], _b200_sa[{{b200_gather_unroll}}];
                    {% for ctype, var, ptr in b200_preloads %}
                    {{ctype}} _b200_rd_{{var}}[{{b200_gather_unroll}}];
                    {% endfor %}
                    #pragma unroll
                    for (int _u = 0; _u < {{b200_gather_unroll}}; ++_u)
                    {
                        const int _slot = _base + 32 * _u + _lane;
                        const bool _sv = _slot < _nslots;
                        const int _sl = _sv ? _slot : 0;
                        int _jlo = 0, _jhi = 32;     // row of the slot: last row with _lexcl <= _sl
                        #pragma unroll
                        for (int _i = 0; _i < 5; ++_i)
                        {
                            const int _mid = (_jlo + _jhi) >> 1;
                            if (__shfl_sync(0xffffffffu, _lexcl, _mid) <= _sl) _jlo = _mid; else _jhi = _mid;
                        }
                        const int _k = __shfl_sync(0xffffffffu, _rbeg, _jlo) + _sl - __shfl_sync(0xffffffffu, _lexcl, _jlo);
                        _b200_sa[_u] = __shfl_sync(0xffffffffu, _srcabs, _jlo);
                        _b200_tg[_u] = _sv ? b200::ld_index(_pw.csr_target + _k) : -1;
                        _b200_sy[_u] = (_sv && !_pw.identity) ? __ldg(_pw.syn_ids + _k) : _k;
                        {% for ctype, var, ptr in b200_preloads %}
                        _b200_rd_{{var}}[_u] = _sv ? {{ptr}}[_b200_sy[_u]] : ({{ctype}})0;
                        {% endfor %}
                    }
                    #pragma unroll
                    for (int _u = 0; _u < {{b200_gather_unroll}}; ++_u)
                    {
                        if (_base + 32 * _u + _lane >= _nslots) break;
                        const int _idx = _b200_sy[_u];
                        const int _b200_tgt_idx = _b200_tg[_u];
                        const int _b200_src_idx = _b200_sa[_u];
                        const int _vectorisation_idx = _idx;
                        {% for ctype, var, ptr in b200_preloads %}
                        const {{ctype}} {{var}} = _b200_rd_{{var}}[_u];
                        {% endfor %}
                        {% if b200_uses_rng %}
                        b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _b200_timestep);
                        {% endif %}
                        {{vector_code|autoindent}}
                    }
                }
            }
            continue;
        }
        long long _a, _b;
        unsigned int _next_ticket = 0u;
        if (_dyn)
        {
            unsigned int _t0 = 0u;
            if (_lane == 0) { _t0 = atomicAdd(_tickets, 1u); _next_ticket = atomicAdd(_tickets, 1u); }
            _t0 = __shfl_sync(0xffffffffu, _t0, 0);
            _a = 32LL * _t0;
            _b = min(_a + 32, _nlines);
        }
        else if (_nrows <= _nwarps)
        {
            // Light step (fewer rows than warps): every warp works on (a part of) ONE row, so the
            // step pays one chain of dependent loads (spike id -> row pointers -> indices), not one
            // per row of a share; `_k` neighbouring warps split the lines of a row.
            const int _k = _nwarps / _nrows;
            const int _r = _gwarp / _k, _part = _gwarp - _r * _k;
            _a = _b = 0;
            if (_r < _nrows)
            {
                const int _L = 31 - __clz(__ballot_sync(0xffffffffu, _rexcl <= _r));
                const int _s = _r - __shfl_sync(0xffffffffu, _rexcl, _L);
                const int _lr = __shfl_sync(0xffffffffu, _blr, _L);
                const int _per = (_lr + _k - 1) / _k;
                const int _lo = min(_lr, _part * _per), _hi = min(_lr, _lo + _per);
                _a = __shfl_sync(0xffffffffu, _bexcl, _L) + (long long)_s * _lr + _lo;
                _b = _a + (_hi - _lo);
            }
        }
        else
        {
            _a = _small ? (long long)(((unsigned int)_gwarp * (unsigned int)_nlines) / (unsigned int)_nwarps)
                        : ((long long)_gwarp * _nlines) / _nwarps;
            _b = _small ? (long long)(((unsigned int)(_gwarp + 1) * (unsigned int)_nlines) / (unsigned int)_nwarps)
                        : ((long long)(_gwarp + 1) * _nlines) / _nwarps;
        }
        for (;;)
        {
        while (_a < _b)
        {
            const int _L = 31 - __clz(__ballot_sync(0xffffffffu, _bexcl <= _a));
            const long long _loc = _a - __shfl_sync(0xffffffffu, _bexcl, _L);
            const int _lr = __shfl_sync(0xffffffffu, _blr, _L);
            const int _delay = __shfl_sync(0xffffffffu, _bdelay, _L);
            const int32_t* _spk = (const int32_t*)__shfl_sync(0xffffffffu, (unsigned long long)_bspk, _L);
            const int* _rp = _pw.rowptr + (size_t)(_g0 + _L) * (_pw.nsrc + 1);
            const int _s = _small ? (int)((unsigned int)_loc / (unsigned int)_lr) : (int)(_loc / _lr);
            const int _l0 = (int)(_loc - (long long)_s * _lr);
            // lines of this row inside my share
            const int _nl = (int)min((long long)(_lr - _l0), _b - _a);
            _a += _nl;
            const int _src = (_delay == _segd ? b200::view_id(_view, _s) : _spk[_s]) - _pw.src_start;
            if (_src < 0 || _src >= _pw.nsrc) continue;
            const int _rbeg = _rp[_src], _rend = _rp[_src + 1];
            if (_l0 == 0) _nev += (unsigned long long)(_rend - _rbeg);
            const int _abeg = (_rbeg & ~31) + 32 * _l0;
            const int _end = min(_rend, _abeg + 32 * _nl);
            const int _b200_src_idx = _src + _pw.src_start;
            int _k0 = _abeg + _lane;
            if (_k0 < _rbeg) _k0 += 32;     // first line of the row: lanes in front of its start
            // {{b200_unroll}} lines of 32 slots per iteration: the loads of the packed index stream(s)
            // are all in flight before the first reduction is issued
            for (int _kb = _k0; _kb < _end; _kb += 32 * {{b200_unroll}})
            {
                int _b200_tg[{{b200_unroll}}], _b200_sy[{{b200_unroll}}];
                {% for ctype, var, ptr in b200_preloads %}
                {{ctype}} _b200_rd_{{var}}[{{b200_unroll}}];
                {% endfor %}
                #pragma unroll
                for (int _u = 0; _u < {{b200_unroll}}; ++_u)
                {
                    const int _k = _kb + 32 * _u;
                    _b200_tg[_u] = _k < _end ? b200::ld_index(_pw.csr_target + _k) : 0;
                    _b200_sy[_u] = (_k < _end && !_pw.identity) ? __ldg(_pw.syn_ids + _k) : _k;
                    {% for ctype, var, ptr in b200_preloads %}
                    _b200_rd_{{var}}[_u] = _k < _end ? {{ptr}}[_b200_sy[_u]] : ({{ctype}})0;
                    {% endfor %}
                }
                #pragma unroll
                for (int _u = 0; _u < {{b200_unroll}}; ++_u)
                {
                    if (_kb + 32 * _u >= _end) break;
                    const int _idx = _b200_sy[_u];
                    const int _b200_tgt_idx = _b200_tg[_u];
                    const int _vectorisation_idx = _idx;
                    {% for ctype, var, ptr in b200_preloads %}
                    const {{ctype}} {{var}} = _b200_rd_{{var}}[_u];
                    {% endfor %}
                    {% if b200_uses_rng %}
                    b200::Rng _rng = b200::rng_init(_A._seed, {{b200_stream_id}}u, _idx, _b200_timestep);
                    {% endif %}
                    {{vector_code|autoindent}}
                }
            }
        }
        if (!_dyn) break;
        // next ticket (requested one ticket ahead: its round trip overlapped the work above)
        const unsigned int _t1 = __shfl_sync(0xffffffffu, _next_ticket, 0);
        _a = 32LL * _t1;
        if (_a >= _nlines) break;
        _b = min(_a + 32, _nlines);
        if (_lane == 0) _next_ticket = atomicAdd(_tickets, 1u);
        }
    }
    // delivered synaptic events (the benchmark metric): one atomic per warp per step
    #pragma unroll
    for (int _o = 16; _o > 0; _o >>= 1) _nev_lane += __shfl_xor_sync(0xffffffffu, _nev_lane, _o);
    _nev += _nev_lane;
    if (_lane == 0 && _nev) atomicAdd(_pw.events, _nev);
    {% endif %}
{% endblock %}

{% block extra_device_code %}
{% if b200_counted %}
// ---- apply pass of the counted pathway: element-private, same partition as the state updater of
// the target group.  The events of this pathway in this step are applied one after the other,
// exactly like the reference's loop over the queue (synapses.cpp:20-49).
__device__ __forceinline__ void _dev_{{codeobj_name}}_apply(const b200::Ctx& _ctx, const _B200Clocks& _clks,
                                                           const _co_{{codeobj_name}}::Scal& _sc)
{
    using namespace _co_{{codeobj_name}};
    ///// CONSTANTS ///////////
    %CONSTANTS_DEV%
    ///// POINTERS ////////////
    {{pointers_lines|autoindent}}
    const b200::PathwayDev& _pw = _A._pw_{{pathway.name}};
    const int64_t _b200_timestep = _clks.{{b200_clock}}.timestep;
    // scalar code
    {{scalar_code|autoindent}}
    if (_pw.tileptr)
    {
        // ---- dense rows: OWNER-COMPUTES in shared memory.  The targets of this CTA (its block of
        // the element partition) are a tile; every (delay bin, source) row was cut at the tile
        // boundaries when the CSR was built (`tileptr`), so this CTA reads exactly the entries of
        // every spiking row that point into its tile -- contiguous pieces of the packed index
        // stream -- and counts them in shared memory: one private counter array per warp, plain
        // read-modify-write (the targets of one row are distinct, a warp handles one row at a
        // time), no atomics of any kind and no global traffic but the index stream itself.
        extern __shared__ int _b200_tile[];
        const b200::EventSpaceDev& _es = {{ '_A._es' + get_array_name(pathway.source.variables[pathway.eventspace_name], access_data=False) }};
        const b200::Slice _mine = b200::owned_cta((int64_t){{b200_counted.size}}, _ctx);
        const int _T = (int)(_mine.hi - _mine.lo);
        const int _lane = threadIdx.x & 31, _warp = threadIdx.x >> 5;
        int* _cnt = _b200_tile + _warp * _pw.tile_stride;
        for (int _k = _lane; _k < _T; _k += 32) _cnt[_k] = 0;
        __syncwarp();
        unsigned long long _nev = 0ULL;
        const int _segd = _pw.seg_delay;
        b200::SpikeView _view;
        _view.total = 0;
        if (_segd >= 0)
            _view = b200::view_build(_es, _b200_timestep - _segd, _ctx, false, _A._ctrl);
        const int _tile = _ctx.gbid;      // tile index = CTA index on this rank
        for (int _bin = 0; _bin < _pw.nbins; ++_bin)
        {
            const int _delay = __ldg(_pw.bin_delay + _bin);
            const int32_t* _spk = 0;
            int _n = _view.total;
            if (_delay != _segd)
            {
                _spk = b200::compact_slot(_es, _b200_timestep - _delay);
                _n = _spk[_es.N];
            }
            const int* _tp = _pw.tileptr + (size_t)_bin * (size_t)(_pw.nsrc + 1) * (size_t)(_ctx.gnb + 1) + _tile;
            // the warp's rows: _s0 + l * kWarps; lane l fetches the descriptor of row l (32
            // chains of dependent loads in flight), then the rows are walked one after the other
            // with the index loads of the next row issued ahead of the counting of this one
            for (int _s0 = _warp; _s0 < _n; _s0 += 32 * b200::kWarps)
            {
                const int _s = _s0 + _lane * b200::kWarps;
                int _beg = 0, _end = 0;
                if (_s < _n)
                {
                    const int _src = (_delay == _segd ? b200::view_id(_view, _s) : _spk[_s]) - _pw.src_start;
                    if (_src >= 0 && _src < _pw.nsrc)
                    {
                        const int* _q = _tp + (size_t)_src * (size_t)(_ctx.gnb + 1);
                        _beg = __ldg(_q);
                        _end = __ldg(_q + 1);
                    }
                }
                _nev += (unsigned long long)(_end - _beg);
                // rows are walked in groups of 4; the index loads of the next group (up to 3 lines
                // of 32 entries per row) are issued before the entries of this group are counted
                const int _nrows = min(32, (_n - _s0 + b200::kWarps - 1) / b200::kWarps);
                int _cur[4][3];
                #pragma unroll
                for (int _r = 0; _r < 4; ++_r)
                {
                    const int _b = __shfl_sync(0xffffffffu, _beg, _r);
                    const int _e = _r < _nrows ? __shfl_sync(0xffffffffu, _end, _r) : _b;
                    #pragma unroll
                    for (int _u = 0; _u < 3; ++_u)
                        _cur[_r][_u] = (_b + 32 * _u + _lane < _e) ? b200::ld_index(_pw.csr_target + _b + 32 * _u + _lane) : -1;
                }
                for (int _j = 0; _j < _nrows; _j += 4)
                {
                    int _nxt[4][3];
                    #pragma unroll
                    for (int _r = 0; _r < 4; ++_r)
                    {
                        const int _jn = _j + 4 + _r;
                        const int _b = __shfl_sync(0xffffffffu, _beg, _jn & 31);
                        const int _e = _jn < _nrows ? __shfl_sync(0xffffffffu, _end, _jn & 31) : _b;
                        #pragma unroll
                        for (int _u = 0; _u < 3; ++_u)
                            _nxt[_r][_u] = (_b + 32 * _u + _lane < _e) ? b200::ld_index(_pw.csr_target + _b + 32 * _u + _lane) : -1;
                    }
                    #pragma unroll
                    for (int _r = 0; _r < 4; ++_r)
                    {
                        {   // the (up to 96) entries of one row point at distinct targets: their
                            // three counter updates are independent (loads first, then stores);
                            // the next row may hit the same counters and has to wait
                            int _old[3];
                            #pragma unroll
                            for (int _u = 0; _u < 3; ++_u)
                                _old[_u] = _cur[_r][_u] >= 0 ? _cnt[_cur[_r][_u] - (int)_mine.lo] : 0;
                            #pragma unroll
                            for (int _u = 0; _u < 3; ++_u)
                                if (_cur[_r][_u] >= 0) _cnt[_cur[_r][_u] - (int)_mine.lo] = _old[_u] + 1;
                            __syncwarp();
                        }
                        // rows with more than 96 entries in this tile: the rest, line by line
                        const int _b = __shfl_sync(0xffffffffu, _beg, (_j + _r) & 31);
                        const int _e = _j + _r < _nrows ? __shfl_sync(0xffffffffu, _end, (_j + _r) & 31) : _b;
                        for (int _k = _b + 96 + _lane; _k - _lane < _e; _k += 32)
                        {
                            if (_k < _e) _cnt[b200::ld_index(_pw.csr_target + _k) - (int)_mine.lo] += 1;
                            __syncwarp();
                        }
                    }
                    #pragma unroll
                    for (int _r = 0; _r < 4; ++_r)
                    {
                        #pragma unroll
                        for (int _u = 0; _u < 3; ++_u) _cur[_r][_u] = _nxt[_r][_u];
                    }
                }
            }
        }
        #pragma unroll
        for (int _o = 16; _o > 0; _o >>= 1) _nev += __shfl_xor_sync(0xffffffffu, _nev, _o);
        if (_lane == 0 && _nev) atomicAdd(_pw.events, _nev);
        __syncthreads();
        for (int _loc = threadIdx.x; _loc < _T; _loc += b200::kBlock)
        {
            int _b200_n = 0;
            #pragma unroll
            for (int _w = 0; _w < b200::kWarps; ++_w) _b200_n += _b200_tile[_w * _pw.tile_stride + _loc];
            if (_b200_n == 0) continue;
            const int _b200_tgt_idx = (int)_mine.lo + _loc;
            const int _idx = _b200_tgt_idx;
            const int _vectorisation_idx = _idx;
            {{b200_apply_loads|autoindent}}
            for (int _b200_k = 0; _b200_k < _b200_n; ++_b200_k)
            {
                {{b200_apply_body|autoindent}}
            }
            {{b200_apply_stores|autoindent}}
        }
        __syncthreads();
        return;
    }
    {% if b200_counted.dual %}
    // (sparse rows of this pathway are delivered by floating-point reductions: nothing to apply)
    {% else %}
    int* _b200_hits = _pw.hits + (size_t)(_b200_timestep % _pw.hits_slots) * (size_t)_pw.hits_n;
    B200_FOR_OWNED(_i64, (int64_t){{b200_counted.size}}, _ctx)
    {
        const int _b200_tgt_idx = (int)_i64;
        const int _b200_n = __ldcg(_b200_hits + _b200_tgt_idx);
        if (_b200_n == 0) continue;
        _b200_hits[_b200_tgt_idx] = 0;      // this slot is counted into again a whole ring later
        const int _idx = _b200_tgt_idx;
        const int _vectorisation_idx = _idx;
        {{b200_apply_loads|autoindent}}
        for (int _b200_k = 0; _b200_k < _b200_n; ++_b200_k)
        {
            {{b200_apply_body|autoindent}}
        }
        {{b200_apply_stores|autoindent}}
    }
    {% endif %}
}

__global__ void __launch_bounds__(b200::kBlock, {{prefs.devices.b200.ctas_per_sm}})
_kernel_{{codeobj_name}}_apply(const _B200Clocks _clks, const _co_{{codeobj_name}}::Scal _sc)
{
    const b200::Ctx _ctx{(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.x, (int)gridDim.x, _A._rank, _A._world};
    b200::view_reset();
    _dev_{{codeobj_name}}_apply(_ctx, _clks, _sc);
}
B200_REGISTER_KERNEL(_kernel_{{codeobj_name}}_apply)

void _run_{{codeobj_name}}_apply()
{
    _co_{{codeobj_name}}::Scal _sc;
    _hostscal_{{codeobj_name}}(_sc);
    _b200_launch_begin("{{codeobj_name}}");
    _kernel_{{codeobj_name}}_apply<<<_b200_grid_size(), b200::kBlock, _b200_dyn_smem, b200::state().stream>>>(_b200_clocks_now(), _sc);
    _b200_launch_end("{{codeobj_name}}");
}
{% endif %}
{% endblock %}
