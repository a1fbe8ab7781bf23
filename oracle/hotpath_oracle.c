/* hotpath_oracle.c -- CPU restatement of the per-timestep hot loop of Brian2's cpp_standalone
 * device.  TEST INFRASTRUCTURE ONLY: nothing under brian2_b200/ may include, link or call this
 * file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against (1) the reference's own
 * known-answer tests for the spike queue (brian2/tests/test_spikequeue.py:36-61) and for delayed
 * delivery (brian2/tests/test_synapses.py:1154-1176), and (2) fixtures produced by running the
 * unmodified reference (cpp_standalone, serial, -ffp-contract=off) on the same inputs
 * (tests/golden/make_oracle_fixtures.py -> tests/golden/oracle_*.npz): LIF networks (CUBA,
 * Brunel with heterogeneous delays), Song-Abbott STDP and the Hodgkin-Huxley network COBAHH --
 * all bit-exact, spike trains and final state.
 *
 * Every function names the reference lines it restates (paths relative to brian2/ in
 * brian-team/brian2).  Plain C11, sequential, no FMA contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* growable int vector                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t* d; int n, cap; } ivec;
static void iv_push(ivec* v, int32_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 16; v->d = (int32_t*)realloc(v->d, sizeof(int32_t) * v->cap); }
    v->d[v->n++] = x;
}
static void iv_free(ivec* v) { free(v->d); v->d = 0; v->n = v->cap = 0; }

/* ------------------------------------------------------------------------------------------ */
/* CSpikeQueue (synapses/spikequeue.h:14-206)                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    ivec* queue;          /* ring of buckets of synapse indices          (spikequeue.h:19)  */
    int qsize;            /* number of buckets = max_delay + 1            (:102)            */
    int offset;           /* current bucket                               (:21)             */
    int* delays;          /* integer delay per synapse                    (:22)             */
    ivec* synapses;       /* synapse indices per source neuron            (:24)             */
    int source_start, source_end;
    int scalar_delay;     /* all delays equal                             (:104)            */
    int n_syn;
} oq_t;

/* CSpikeQueue::prepare, spikequeue.h:48-105 (first call: no re-scaling of stored spikes). */
oq_t* oq_create(const double* real_delays, int n_delays, const int32_t* sources, int n_syn,
                double dt, int source_start, int source_end) {
    oq_t* q = (oq_t*)calloc(1, sizeof(oq_t));
    q->source_start = source_start; q->source_end = source_end; q->n_syn = n_syn;
    const int nsrc = source_end - source_start;
    q->synapses = (ivec*)calloc(nsrc > 0 ? nsrc : 1, sizeof(ivec));
    q->delays = (int*)calloc(n_syn > 0 ? n_syn : 1, sizeof(int));
    int max_delay = 0, all_equal = 1;
    for (int i = 0; i < n_syn; ++i) {
        const double rd = n_delays == 1 ? real_delays[0] : real_delays[i];
        q->delays[i] = (int)(rd / dt + 0.5);                         /* :90 */
        if (q->delays[i] > max_delay) max_delay = q->delays[i];
        if (q->delays[i] != q->delays[0]) all_equal = 0;
        iv_push(&q->synapses[sources[i] - source_start], i);        /* :96-97 */
    }
    q->scalar_delay = all_equal;                                     /* :104 (n_delays==1 or equal) */
    q->qsize = max_delay + 1;                                        /* :102 */
    q->queue = (ivec*)calloc(q->qsize, sizeof(ivec));
    q->offset = 0;
    return q;
}
void oq_destroy(oq_t* q) {
    for (int i = 0; i < q->qsize; ++i) iv_free(&q->queue[i]);
    for (int i = 0; i < q->source_end - q->source_start; ++i) iv_free(&q->synapses[i]);
    free(q->queue); free(q->synapses); free(q->delays); free(q);
}
/* CSpikeQueue::push, spikequeue.h:151-191: spikes is ascending; restrict to the source range,
 * append each spiking neuron's synapses to bucket (offset + delay) % size. */
void oq_push(oq_t* q, const int32_t* spikes, int nspikes) {
    int lo = 0, hi = nspikes;
    while (lo < nspikes && spikes[lo] < q->source_start) ++lo;      /* lower_bound :154 */
    hi = lo;
    while (hi < nspikes && spikes[hi] < q->source_end) ++hi;        /* upper_bound :155 */
    for (int s = lo; s < hi; ++s) {
        const ivec* syn = &q->synapses[spikes[s] - q->source_start];
        for (int k = 0; k < syn->n; ++k) {
            const int id = syn->d[k];
            iv_push(&q->queue[(q->offset + q->delays[id]) % q->qsize], id);
        }
    }
}
/* peek (:193-196) */
const int32_t* oq_peek(const oq_t* q, int* n) { *n = q->queue[q->offset].n; return q->queue[q->offset].d; }
/* advance (:198-205): clear the current bucket, move on */
void oq_advance(oq_t* q) { q->queue[q->offset].n = 0; q->offset = (q->offset + 1) % q->qsize; }

/* ------------------------------------------------------------------------------------------ */
/* helper functions of the C++ target (codegen/generators/cpp_generator.py)                    */
/* ------------------------------------------------------------------------------------------ */
static int64_t o_timestep(double t, double dt) { return (int64_t)((t + 1e-3 * dt) / dt); }   /* :651-656 */
static double o_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); } /* :627-640 */

/* ------------------------------------------------------------------------------------------ */
/* Pathway description shared by the network drivers                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const int32_t* pre;      /* _synaptic_pre  (absolute index in the parent group)            */
    const int32_t* post;     /* _synaptic_post                                                 */
    const double* delay;     /* n_delays entries (1 = homogeneous)                             */
    int n_delays;
    int n_syn;
    int source_start, source_end;
    int target_var;          /* 0: v, 1: ge, 2: gi                                             */
    double weight;           /* scalar weight (used when w == NULL)                            */
    const double* w;         /* optional per-synapse weights                                   */
} opath_t;

typedef struct {
    /* per-step propagator of the LIF family, exactly as the reference's generated vector code:
     *   not_refractory = _timestep(t - lastspike, dt) >= ref_steps
     *   _ge = a_ge*ge; _gi = a_gi*gi
     *   _v  = not_refractory ? c0 + (((c_ge*ge) + (c_gi*gi)) + (c_v*v)) : c_ref + v
     *   ge=_ge; gi=_gi; if(not_refractory) v=_v                                             */
    int64_t ref_steps;
    double a_ge, a_gi, c_ref, c0, c_ge, c_gi, c_v;
    double v_thresh, v_reset;
    int use_ge_gi;           /* 0: the model has no ge/gi (Brunel): v = c0 + (c_v*v)           */
} olif_t;

/* Network::run (devices/cpp_standalone/templates/network.cpp:38-122) for one NeuronGroup of the
 * LIF family + P pathways + SpikeMonitor + PopulationRateMonitor.  Slot order per step:
 * stateupdate (templates/stateupdate.cpp:5-22), threshold (threshold.cpp:3-37), monitors
 * (spikemonitor.cpp:6-51, ratemonitor.cpp:6-36), then for each pathway in the given order
 * advance+push (synapses_push_spikes.cpp:26-27) and the effect loop (synapses.cpp:11-50), then
 * reset (reset.cpp:3-23); finally Clock::tick (brianlib/clocks.h:34-38).
 * Outputs: spike monitor (i, t) up to cap_spikes entries, count per neuron, rate per step.
 * Returns the number of recorded spikes (may exceed cap_spikes: caller must re-run larger). */
long long oracle_lif_run(int N, double* v, double* ge, double* gi, double* lastspike,
                         char* not_refractory, const olif_t* m, const opath_t* paths, int n_paths,
                         double dt, long long n_steps, int32_t* mon_i, double* mon_t,
                         long long cap_spikes, int32_t* mon_count, double* rate,
                         double* events_out) {
    oq_t** queues = (oq_t**)calloc(n_paths > 0 ? n_paths : 1, sizeof(oq_t*));
    for (int p = 0; p < n_paths; ++p)
        queues[p] = oq_create(paths[p].delay, paths[p].n_delays, paths[p].pre, paths[p].n_syn, dt,
                              paths[p].source_start, paths[p].source_end);
    int32_t* spikespace = (int32_t*)calloc((size_t)N + 1, sizeof(int32_t));
    long long n_rec = 0;
    double events = 0.0;
    for (long long timestep = 0; timestep < n_steps; ++timestep) {
        const double t = timestep * dt;                               /* clocks.h:37 */
        /* ---- stateupdate ---- */
        for (int i = 0; i < N; ++i) {
            const char nr = o_timestep(t - lastspike[i], dt) >= m->ref_steps;
            double _v;
            if (m->use_ge_gi) {
                const double _ge = m->a_ge * ge[i];
                const double _gi = m->a_gi * gi[i];
                if (!nr) _v = m->c_ref + v[i];
                else _v = m->c0 + (((m->c_ge * ge[i]) + (m->c_gi * gi[i])) + (m->c_v * v[i]));
                ge[i] = _ge; gi[i] = _gi;
            } else {
                if (!nr) _v = m->c_ref + v[i];
                else _v = m->c0 + (m->c_v * v[i]);
            }
            if (nr) v[i] = _v;
            not_refractory[i] = nr;
        }
        /* ---- threshold ---- */
        int count = 0;
        for (int i = 0; i < N; ++i) {
            const char cond = not_refractory[i] ? (v[i] > m->v_thresh) : 0;
            if (cond) { spikespace[count++] = i; not_refractory[i] = 0; lastspike[i] = t; }
        }
        spikespace[N] = count;
        /* ---- monitors ---- */
        for (int s = 0; s < count; ++s) {
            if (n_rec < cap_spikes) { mon_i[n_rec] = spikespace[s]; mon_t[n_rec] = t; }
            ++n_rec;
            if (mon_count) mon_count[spikespace[s]]++;
        }
        if (rate) rate[timestep] = 1.0 * count / dt / N;               /* ratemonitor.cpp:33 */
        /* ---- pathways ---- */
        for (int p = 0; p < n_paths; ++p) {
            oq_t* q = queues[p];
            oq_advance(q);
            oq_push(q, spikespace, count);
            int ns; const int32_t* ids = oq_peek(q, &ns);
            events += ns;
            double* target = paths[p].target_var == 0 ? v : (paths[p].target_var == 1 ? ge : gi);
            for (int k = 0; k < ns; ++k) {
                const int id = ids[k];
                const int j = paths[p].post[id];
                /* `v` carries the "(unless refractory)" flag: synaptic writes to it are
                 * conditional on not_refractory of the target (conditional write,
                 * codegen/generators/cpp_generator.py:287-314) */
                if (paths[p].target_var == 0 && !not_refractory[j]) continue;
                target[j] += paths[p].w ? paths[p].w[id] : paths[p].weight;
            }
        }
        /* ---- reset ---- */
        for (int s = 0; s < count; ++s) v[spikespace[s]] = m->v_reset;
    }
    for (int p = 0; p < n_paths; ++p) oq_destroy(queues[p]);
    free(queues); free(spikespace);
    if (events_out) *events_out = events;
    return n_rec;
}

/* Same driver written so that python can pass the pathways as flat arrays. */
long long oracle_lif_run_flat(int N, double* v, double* ge, double* gi, double* lastspike,
                              char* not_refractory, const double* mpar, int use_ge_gi,
                              int n_paths, const int32_t** pre, const int32_t** post,
                              const double** delay, const int* n_delays, const int* n_syn,
                              const int* src_start, const int* src_end, const int* target_var,
                              const double* weight, const double** w, double dt, long long n_steps,
                              int32_t* mon_i, double* mon_t, long long cap_spikes,
                              int32_t* mon_count, double* rate, double* events_out) {
    olif_t m;
    m.ref_steps = (int64_t)mpar[0]; m.a_ge = mpar[1]; m.a_gi = mpar[2]; m.c_ref = mpar[3];
    m.c0 = mpar[4]; m.c_ge = mpar[5]; m.c_gi = mpar[6]; m.c_v = mpar[7];
    m.v_thresh = mpar[8]; m.v_reset = mpar[9]; m.use_ge_gi = use_ge_gi;
    opath_t* paths = (opath_t*)calloc(n_paths > 0 ? n_paths : 1, sizeof(opath_t));
    for (int p = 0; p < n_paths; ++p) {
        paths[p].pre = pre[p]; paths[p].post = post[p]; paths[p].delay = delay[p];
        paths[p].n_delays = n_delays[p]; paths[p].n_syn = n_syn[p];
        paths[p].source_start = src_start[p]; paths[p].source_end = src_end[p];
        paths[p].target_var = target_var[p]; paths[p].weight = weight[p];
        paths[p].w = w ? w[p] : 0;
    }
    const long long r = oracle_lif_run(N, v, ge, gi, lastspike, not_refractory, &m, paths, n_paths,
                                       dt, n_steps, mon_i, mon_t, cap_spikes, mon_count, rate,
                                       events_out);
    free(paths);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* Song-Abbott STDP (examples/synapses/STDP.py:30-62) as generated by the reference:           */
/*   inputs:  x += dt*rate ; spike if x > 1 ; reset x = 0                                      */
/*   neuron:  _ge = (l1*ge)+ge ; _v = (l2*((El + (ge*(Ee - v))) - v)) + v ; spike v > vt      */
/*   on_pre / on_post bodies: see SURVEY.md App. A.5 (templates/synapses.cpp:39-46 executes    */
/*   them sequentially in queue order)                                                         */
/* par: taue, taum, El, Ee, vt, vr, taupre, taupost, dApre, dApost, gmax                       */
/* ------------------------------------------------------------------------------------------ */
long long oracle_stdp_run(int N_in, double* x, const double* rate, double* v1, double* ge1,
                          double* w, double* Apre, double* Apost, double* lastupdate,
                          const double* par, double dt, long long n_steps, int32_t* in_i,
                          double* in_t, long long cap_in, double* out_t, long long cap_out,
                          long long* n_out_spikes) {
    const double taue = par[0], taum = par[1], El = par[2], Ee = par[3], vt = par[4], vr = par[5],
                 taupre = par[6], taupost = par[7], dApre = par[8], dApost = par[9], gmax = par[10];
    /* all-to-one connectivity: synapse k = (pre k, post 0), no delays (S.connect()) */
    int32_t* pre = (int32_t*)malloc(sizeof(int32_t) * N_in);
    int32_t* post = (int32_t*)calloc(N_in, sizeof(int32_t));
    for (int k = 0; k < N_in; ++k) pre[k] = k;
    const double zero = 0.0;
    oq_t* qpre = oq_create(&zero, 1, pre, N_in, dt, 0, N_in);
    oq_t* qpost = oq_create(&zero, 1, post, N_in, dt, 0, 1);
    int32_t* ss_in = (int32_t*)calloc((size_t)N_in + 1, sizeof(int32_t));
    int32_t ss_out[2] = {0, 0};
    long long n_in = 0, n_out = 0;
    double v = v1[0], ge = ge1[0];
    const double l1n = 1.0f * (-dt) / taue, l2n = 1.0f * dt / taum;
    const double lpost = 1.0f * 1.0 / taupost, lpre = 1.0f * 1.0 / taupre;
    for (long long timestep = 0; timestep < n_steps; ++timestep) {
        const double t = timestep * dt;
        for (int i = 0; i < N_in; ++i) x[i] = (dt * rate[i]) + x[i];
        {
            const double _ge = (l1n * ge) + ge;
            const double _v = (l2n * ((El + (ge * (Ee - v))) - v)) + v;
            ge = _ge; v = _v;
        }
        int c_in = 0;
        for (int i = 0; i < N_in; ++i) if (x[i] > 1) ss_in[c_in++] = i;
        ss_in[N_in] = c_in;
        int c_out = 0;
        if (v > vt) ss_out[c_out++] = 0;
        ss_out[1] = c_out;
        for (int s = 0; s < c_in; ++s) { if (n_in < cap_in) { in_i[n_in] = ss_in[s]; in_t[n_in] = t; } ++n_in; }
        if (c_out) { if (n_out < cap_out) out_t[n_out] = t; ++n_out; }
        /* on_pre */
        oq_advance(qpre); oq_push(qpre, ss_in, c_in);
        { int ns; const int32_t* ids = oq_peek(qpre, &ns);
          for (int k = 0; k < ns; ++k) {
              const int id = ids[k];
              const double _Apost = Apost[id] * exp(lpost * (-(t - lastupdate[id])));
              const double _Apre = Apre[id] * exp(lpre * (-(t - lastupdate[id])));
              Apost[id] = _Apost; Apre[id] = _Apre;
              ge += w[id];
              Apre[id] += dApre;
              w[id] = o_clip(w[id] + Apost[id], 0, gmax);
              lastupdate[id] = t;
          } }
        /* on_post */
        oq_advance(qpost); oq_push(qpost, ss_out, c_out);
        { int ns; const int32_t* ids = oq_peek(qpost, &ns);
          for (int k = 0; k < ns; ++k) {
              const int id = ids[k];
              const double _Apost = Apost[id] * exp(lpost * (-(t - lastupdate[id])));
              const double _Apre = Apre[id] * exp(lpre * (-(t - lastupdate[id])));
              Apost[id] = _Apost; Apre[id] = _Apre;
              Apost[id] += dApost;
              w[id] = o_clip(w[id] + Apre[id], 0, gmax);
              lastupdate[id] = t;
          } }
        for (int s = 0; s < c_in; ++s) x[ss_in[s]] = 0;
        if (c_out) v = vr;
    }
    v1[0] = v; ge1[0] = ge;
    *n_out_spikes = n_out;
    oq_destroy(qpre); oq_destroy(qpost);
    free(pre); free(post); free(ss_in);
    return n_in;
}

/* ------------------------------------------------------------------------------------------ */
/* COBAHH (examples/COBAHH.py:44-59 equations; BASELINE.json configs[1])                       */
/* ------------------------------------------------------------------------------------------ */
/* _exprel, codegen/generators/cpp_generator.py:603-611 */
static double o_exprel(double x) {
    if (fabs(x) < 1e-16) return 1.0;
    if (x > 717) return INFINITY;
    return expm1(x) / x;
}

/* One run of the Hodgkin-Huxley network of tests/models.py:cobahh: exponential-Euler state update
 * (stateupdaters/exponential_euler.py:82: x <- -B/A + (B/A + x) exp(A dt) per variable, all from
 * the OLD state; the grouping of the terms below is the one of the abstract code the reference
 * generates for these equations -- rounding follows it), refractory threshold `v > -20 mV`
 * without reset (templates/threshold.cpp:3-37 with `not_refractory`, `lastspike`), two delay-free
 * pathways `ge += we` (sources [0, Ne)) and `gi += wi` (sources [Ne, N)) delivered in this order
 * in the step of the spike (synapses_push_spikes.cpp:26-27, synapses.cpp:11-50), SpikeMonitor.
 * par = { Cm, gl, El, EK, ENa, g_na, g_kd, VT, taue, taui, Ee, Ei, we, wi, refractory }. */
long long oracle_hh_run(int N, int Ne, double* v, double* ge, double* gi, double* m, double* n,
                        double* h, double* lastspike, char* not_refractory, const double* par,
                        const int32_t* ce_pre, const int32_t* ce_post, int n_ce,
                        const int32_t* ci_pre, const int32_t* ci_post, int n_ci,
                        double dt, long long n_steps, int32_t* mon_i, double* mon_t,
                        long long cap_spikes, int32_t* count, double* events_out) {
    const double Cm = par[0], gl = par[1], El = par[2], EK = par[3], ENa = par[4], g_na = par[5],
                 g_kd = par[6], VT = par[7], taue = par[8], taui = par[9], Ee = par[10], Ei = par[11],
                 we = par[12], wi = par[13], refractory = par[14];
    const double ms = 0.001, mV = 0.001;
    /* loop-invariant scalars hoisted by codegen/optimisation.py (evaluated once per step there) */
    const int64_t ref_steps = o_timestep(refractory, dt);
    const double a_ge = exp(((-1.0) * dt) / taue), a_gi = exp(((-1.0) * dt) / taui);
    const double k_ah = (0.128 * (2.5713844347880297 * exp((0.05555555555555555 * VT) / mV))) / ms;
    const double s18 = 0.05555555555555555 / mV;
    const double k_bh = (ms * 2980.9579870417283) * exp((0.2 * VT) / mV);
    const double s5 = 0.2 / mV;
    const double k_ah2 = (0.128 * (2.5713844347880297 * pow(exp(VT / mV), 0.05555555555555555))) / ms;
    const double s1 = 1.0 / mV;
    const double k_m = 1.28 / ms, k_mneg = (-1.28) / ms;
    const double o_am = ((0.25 * VT) / mV) + 3.25, s4 = 0.25 / mV;
    const double k_bm = 1.4 / ms;
    const double o_bm = (-8.0) + ((0.2 * (-VT)) / mV);
    const double k_an = 0.16 / ms;
    const double k_bn = ((-0.6420127083438707) * pow(exp(VT / mV), 0.025)) / ms;
    const double o_an = 3.0 + ((0.2 * VT) / mV);
    const double c_l = (El * gl) / Cm, c_k = (EK * g_kd) / Cm, c_na = (ENa * g_na) / Cm;
    const double c_e = Ee / Cm, c_i = Ei / Cm;
    const double a_l = 0.0 - (gl / Cm), a_k = (-g_kd) / Cm, a_na = g_na / Cm, a_s = 1.0 / Cm;
    /* per-source synapse lists in synapse-index order (CSpikeQueue::prepare, spikequeue.h:96-97) */
    int* ce_ptr = (int*)calloc((size_t)N + 2, sizeof(int));
    int* ci_ptr = (int*)calloc((size_t)N + 2, sizeof(int));
    for (int k = 0; k < n_ce; ++k) ce_ptr[ce_pre[k] + 1]++;
    for (int k = 0; k < n_ci; ++k) ci_ptr[ci_pre[k] + 1]++;
    for (int i = 0; i < N; ++i) { ce_ptr[i + 1] += ce_ptr[i]; ci_ptr[i + 1] += ci_ptr[i]; }
    int* ce_syn = (int*)malloc(sizeof(int) * (size_t)(n_ce > 0 ? n_ce : 1));
    int* ci_syn = (int*)malloc(sizeof(int) * (size_t)(n_ci > 0 ? n_ci : 1));
    {
        int* cur = (int*)malloc(sizeof(int) * ((size_t)N + 1));
        memcpy(cur, ce_ptr, sizeof(int) * ((size_t)N + 1));
        for (int k = 0; k < n_ce; ++k) ce_syn[cur[ce_pre[k]]++] = k;
        memcpy(cur, ci_ptr, sizeof(int) * ((size_t)N + 1));
        for (int k = 0; k < n_ci; ++k) ci_syn[cur[ci_pre[k]]++] = k;
        free(cur);
    }
    int32_t* spikes = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    long long n_rec = 0;
    double events = 0.0;
    for (long long step = 0; step < n_steps; ++step) {
        const double t = (double)step * dt;                           /* clocks.h:34-38 */
        for (int i = 0; i < N; ++i) {                                 /* stateupdate.cpp:5-22 */
            const double vi = v[i], gei = ge[i], gii = gi[i], mi = m[i], ni = n[i], hi = h[i];
            not_refractory[i] = o_timestep(t - lastspike[i], dt) >= ref_steps;
            const double new_ge = a_ge * gei, new_gi = a_gi * gii;
            /* h:  A = -(alpha_h + beta_h), B/A as generated */
            const double e5 = exp(s5 * (-vi));
            const double p18 = pow(exp(s1 * vi), 0.05555555555555555);
            const double A_h = ((-4.0) / (ms + (k_bh * e5))) - (k_ah2 / p18);
            const double BA_h = (k_ah * exp(s18 * (-vi))) / A_h;
            const double new_h = (-BA_h) + ((BA_h + hi) * exp(dt * A_h));
            /* m */
            const double x_am = o_exprel(o_am - (s4 * vi)), x_bm = o_exprel(o_bm + (s5 * vi));
            const double A_m = (k_mneg / x_am) - (k_bm / x_bm);
            const double BA_m = k_m / (A_m * x_am);
            const double new_m = (-BA_m) + ((BA_m + mi) * exp(dt * A_m));
            /* n */
            const double p40 = pow(exp(s1 * vi), 0.025), x_an = o_exprel(o_an - (s5 * vi));
            const double A_n = (k_bn / p40) - (k_an / x_an);
            const double BA_n = k_an / (A_n * x_an);
            const double new_n = (-BA_n) + ((BA_n + ni) * exp(dt * A_n));
            /* v */
            const double n4 = pow(ni, 4), m3 = pow(mi, 3);
            const double A_v = (a_l + (a_k * n4)) - (((a_na * (hi * m3)) + (a_s * gei)) + (a_s * gii));
            const double B_v = c_l + ((((c_k * n4) + (c_na * (hi * m3))) + (c_e * gei)) + (c_i * gii));
            const double BA_v = B_v / A_v;
            const double new_v = (-BA_v) + ((BA_v + vi) * exp(dt * A_v));
            ge[i] = new_ge; gi[i] = new_gi; h[i] = new_h; m[i] = new_m; n[i] = new_n; v[i] = new_v;
        }
        int ns = 0;                                                   /* threshold.cpp:18-31 */
        for (int i = 0; i < N; ++i) {
            const int cond = not_refractory[i] ? (v[i] > (-20.0) * mV) : 0;
            if (cond) { spikes[ns++] = i; not_refractory[i] = 0; lastspike[i] = t; }
        }
        for (int s = 0; s < ns; ++s) {                                /* spikemonitor.cpp:6-51 */
            if (n_rec < cap_spikes) { mon_i[n_rec] = spikes[s]; mon_t[n_rec] = t; }
            n_rec++;
            count[spikes[s]]++;
        }
        for (int s = 0; s < ns; ++s) {                                /* hh_Ce: ge += we */
            const int src = spikes[s];
            if (src >= Ne) continue;
            for (int k = ce_ptr[src]; k < ce_ptr[src + 1]; ++k) { ge[ce_post[ce_syn[k]]] += we; events += 1.0; }
        }
        for (int s = 0; s < ns; ++s) {                                /* hh_Ci: gi += wi */
            const int src = spikes[s];
            if (src < Ne) continue;
            for (int k = ci_ptr[src]; k < ci_ptr[src + 1]; ++k) { gi[ci_post[ci_syn[k]]] += wi; events += 1.0; }
        }
    }
    if (events_out) *events_out = events;
    free(ce_ptr); free(ci_ptr); free(ce_syn); free(ci_syn); free(spikes);
    return n_rec;
}
