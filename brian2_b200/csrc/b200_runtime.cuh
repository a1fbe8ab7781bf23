// b200_runtime.cuh -- device-side runtime of the Brian2 `b200` simulation device (sm_100a).
//
// Hand-written building blocks that the generated code objects (templates/*.cu) are inlined
// into.  Everything here is model independent; the per-model abstract code arrives as the
// body of small __device__ functions.
//
// Reference semantics this replaces (brian-team/brian2, all under brian2/):
//   * time loop / clock tick ........ devices/cpp_standalone/templates/network.cpp:38-122,
//                                     devices/cpp_standalone/brianlib/clocks.h:34-38
//   * ordered spike compaction ...... devices/cpp_standalone/templates/threshold.cpp:18-31
//   * spike queue (delay buffer) .... synapses/spikequeue.h:14-206 (CSpikeQueue),
//                                     devices/cpp_standalone/templates/synapses_classes.cpp:15-88
//   * synaptic propagation loop ..... devices/cpp_standalone/templates/synapses.cpp:20-49
//
// Design (B200 first, not a translation):
//   * Work partition.  Every per-element loop (state update, threshold) uses ONE mapping from
//     element index to (CTA, warp, lane): a CTA owns a contiguous chunk (multiple of 512), a
//     warp owns a contiguous slice of it (multiple of 32) and strides through it lane-wise, so
//     a warp load is one fully coalesced 256-byte (fp64) request and, because the mapping is
//     identical in every code object, element-private read-after-write chains between code
//     objects (stateupdate -> threshold) need no grid barrier at all.
//   * Spike "queue".  There is no queue of synapse ids.  Every event space is a ring of past
//     spike lists (`ring_slots` x (N+1) int32, slot = timestep % ring_slots) that the
//     thresholder writes IN PLACE, so a push costs zero bytes.  Delays are resolved at
//     delivery: synapses are stored CSR-by-(delay bin, source neuron) and a pathway visits, for
//     every delay bin d, the rows of the neurons that spiked d steps ago.
//   * Ordered compaction.  Warp ballots + one intra-CTA scan + a single-pass decoupled
//     look-back across CTAs (all CTAs are co-resident by construction), giving the same
//     ascending `_spikespace` the serial reference loop produces.
//   * The whole step loop can run inside one persistent cooperative kernel; code objects are
//     separated by a hand-rolled grid barrier only where the generator's read/write-set
//     analysis finds a cross-thread dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "b200_types.h"

namespace b200 {


// ---------------------------------------------------------------------------------------------
// memory-model helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p) {
    long long v;
    asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One word of the packed CSR index stream (immutable during a run: non-coherent load).  With
// -DB200_CSR_EVICT_LAST (prefs.devices.b200.csr_l2_evict_last) the line is tagged evict_last in L2.
__device__ __forceinline__ int ld_index(const int* p) {
#ifdef B200_CSR_EVICT_LAST
    unsigned long long pol;
    int v;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}

// ---------------------------------------------------------------------------------------------
// system-scope flavours (multi-GPU: data and flags cross NVLink into peer memory)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Tags: every word of a segment names the time step it belongs to (never 0, so a
// zero-initialised slot is never mistaken for a step): counts carry step_tag << 16, ids id_tag << 32.
__host__ __device__ __forceinline__ int step_tag(int64_t timestep) {
    int64_t m = (timestep + 1) % 32767;
    if (m < 0) m += 32767;
    return (int)m + 1;
}
__host__ __device__ __forceinline__ unsigned long long id_tag(int64_t timestep) {
    return ((unsigned long long)(unsigned int)(timestep + 1)) << 32;
}

// ---------------------------------------------------------------------------------------------
// Grid barrier for the persistent kernel.  Monotonic 64-bit arrival counter (never reset inside
// a launch); `target` is thread-0 private state.  Release on arrival, acquire fence after the wait
// (the gpu-scope fence also invalidates this SM's L1, so plain loads after the barrier see other
// CTAs' writes).  Bit 62 of the counter word is the "leave the step loop" flag: whoever raises
// it does so before its own arrival, so the value a poller finally sees carries the flag and the
// end-of-step check costs no extra round trip to L2.  Returns the flag (to all threads).
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kStopBit = 1ULL << 62;

__device__ __forceinline__ void raise_stop(Control* ctrl) {
    ctrl->stop = 1;
    atomicOr(&ctrl->barrier, kStopBit);
}

__device__ __forceinline__ bool grid_barrier(unsigned long long* counter, unsigned long long& target,
                                             const Ctx& c) {
    __shared__ int s_flag;
    __syncthreads();
    if (threadIdx.x == 0) {
        target += (unsigned long long)c.gnb;
        // arrive: release-RMW without a return value (the CTA's earlier writes are ordered before
        // it through the bar.sync above; no round trip to wait for before polling starts)
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ULL) : "memory");
        // wait: relaxed polls (served by L2, no L1 invalidation per iteration), one acquire
        // fence at the end -- it also drops this SM's stale L1 lines for the plain loads that follow
        unsigned long long v;
        while (((v = ld_relaxed_u64(counter)) & ~kStopBit) < target) {
        }
        s_flag = (v & kStopBit) ? 1 : 0;
        __threadfence();
    }
    __syncthreads();
    return s_flag != 0;
}

// ---------------------------------------------------------------------------------------------
// Partition of the elements of a group: first over ranks (contiguous blocks, multiple of 32),
// then over the warps of the grid in units of 32-element "tasks" (contiguous per warp, so the
// spike ids a CTA produces are ascending and the ids of CTA b precede those of CTA b+1).
// Every per-element code object uses this one mapping, so element-private read-after-write
// chains between code objects (stateupdate -> threshold -> reset) need no grid barrier.
// ---------------------------------------------------------------------------------------------
struct Slice {
    int64_t lo;   // first element of this warp's slice
    int64_t hi;   // one past the last element
};

__host__ __device__ __forceinline__ void rank_range(int64_t N, int rank, int world, int64_t& lo, int64_t& hi) {
    if (world == 1) { lo = 0; hi = N; return; }
    int64_t per = (N + world - 1) / world;
    per = (per + 31) & ~(int64_t)31;
    lo = (int64_t)rank * per; if (lo > N) lo = N;
    hi = lo + per; if (hi > N) hi = N;
}
// first element owned by global warp g (of G) inside the rank range [lo, hi)
__host__ __device__ __forceinline__ int64_t warp_first(int64_t lo, int64_t hi, int64_t g, int64_t G) {
    const int64_t T = (hi - lo + 31) >> 5;
    // (same quotient either way; the 32-bit division is ~10x cheaper on the device and covers
    // every group of fewer than 2^32 / (32 G) elements per rank)
    const int64_t q = (T * G < 0xffffffffLL) ? (int64_t)((unsigned int)(g * T) / (unsigned int)G)
                                              : (g * T) / G;
    int64_t e = lo + 32 * q;
    return e < hi ? e : hi;
}

// The slices of a launch never change, and the division above is ~100 instructions: a small
// per-CTA cache in shared memory keeps the first element of every warp of the CTA for the last
// few group sizes seen.  COLLECTIVE: every thread of the CTA must call owned_slice/owned_cta at
// the same point with the same N (true for all per-element templates); a miss synchronises.
constexpr int kSliceCache = 4;
struct SliceCache {
    long long N[kSliceCache];
    long long first[kSliceCache][kWarps + 1];
    int next;
};
__device__ __forceinline__ SliceCache& slice_cache() {
    __shared__ SliceCache sc;
    return sc;
}
// to be called once at kernel start (before the first __syncthreads)
__device__ __forceinline__ void slice_cache_reset() {
    SliceCache& sc = slice_cache();
    if (threadIdx.x < kSliceCache) sc.N[threadIdx.x] = -1;
    if (threadIdx.x == 0) sc.next = 0;
}
__device__ __forceinline__ int slice_cache_slot(int64_t N, const Ctx& c) {
    SliceCache& sc = slice_cache();
    int slot = -1;
#pragma unroll
    for (int k = 0; k < kSliceCache; ++k)
        if (sc.N[k] == (long long)N) slot = k;
    if (slot < 0) {             // uniform over the CTA
        __syncthreads();
        slot = sc.next;
        if (threadIdx.x <= kWarps) {
            int64_t lo, hi;
            rank_range(N, c.rank, c.world, lo, hi);
            sc.first[slot][threadIdx.x] = warp_first(lo, hi, (int64_t)c.gbid * kWarps + threadIdx.x,
                                                     (int64_t)c.gnb * kWarps);
        }
        __syncthreads();
        if (threadIdx.x == 0) { sc.N[slot] = (long long)N; sc.next = (slot + 1) % kSliceCache; }
        __syncthreads();
    }
    return slot;
}

__device__ __forceinline__ Slice owned_slice(int64_t N, const Ctx& c) {
    const int slot = slice_cache_slot(N, c);
    const SliceCache& sc = slice_cache();
    Slice s;
    s.lo = sc.first[slot][threadIdx.x >> 5];
    s.hi = sc.first[slot][(threadIdx.x >> 5) + 1];
    return s;
}
// the elements owned by the whole CTA
__device__ __forceinline__ Slice owned_cta(int64_t N, const Ctx& c) {
    const int slot = slice_cache_slot(N, c);
    const SliceCache& sc = slice_cache();
    Slice s;
    s.lo = sc.first[slot][0];
    s.hi = sc.first[slot][kWarps];
    return s;
}

// for (idx in my lane's elements)
#define B200_FOR_OWNED(IDX, N_, CTX)                                                        \
    const b200::Slice _b200_sl = b200::owned_slice((int64_t)(N_), (CTX));                    \
    for (int64_t IDX = _b200_sl.lo + (threadIdx.x & 31); IDX < _b200_sl.hi; IDX += 32)

__device__ __forceinline__ int ring_index(int64_t timestep, int slots) {
    int64_t s = timestep % slots;
    if (s < 0) s += slots;
    return (int)s;
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// Thresholder output: the CTA's spikes go, ascending, into the CTA's own segment of the event
// space -- no communication with other CTAs (the reference's loop is serial, threshold.cpp:18-31).
//   mask   bit k of lane l  <=>  element (slice.lo + 32*k + l) fired
// Called by ALL threads of ALL CTAs.  On several GPUs every store is repeated into each peer's
// ring (NVLink P2P).  No ordering between the stores is needed: every word (count and id alike)
// carries the tag of its time step, and a reader spins on exactly the word it needs until the
// tag matches (view_build for counts, view_load for ids).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void publish_owned(unsigned long long mask, int niter, const Ctx& c,
                                              const EventSpaceDev& es, int64_t timestep) {
    __shared__ int s_warp[kWarps];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const Slice sl = owned_slice(es.N, c);
    const size_t slot = (size_t)ring_index(timestep, es.slots);
    const int segi = c.rank * c.gnb + c.gbid;

    int wtotal = __popcll(mask);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wtotal += __shfl_xor_sync(0xffffffffu, wtotal, o);
    __syncthreads();            // s_warp may still be read by a previous call
    if (lane == 0) s_warp[warp] = wtotal;
    __syncthreads();
    int woff = 0, ctotal = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int v = s_warp[w];
        if (w < warp) woff += v;
        ctotal += v;
    }
    if (threadIdx.x == 0) {
        const int word = (step_tag(timestep) << 16) | ctotal;
        const size_t o = slot * (size_t)es.nseg + segi;
        es.cnt[o] = word;
        for (int q = 0; q < c.world; ++q)
            if (q != c.rank) es.peer_cnt[q][o] = word;
    }
    if (wtotal > 0) {
        const size_t base = slot * (size_t)es.N + (size_t)es.seg_start[segi];
        const unsigned long long tag = id_tag(timestep);
        int pos = woff;
        for (int k = 0; k < niter; ++k) {
            const bool f = (mask >> k) & 1ULL;
            const unsigned int bal = __ballot_sync(0xffffffffu, f);
            if (f) {
                const int p = pos + __popc(bal & ((1u << lane) - 1u));
                const unsigned long long w = tag | (unsigned int)(sl.lo + 32 * (int64_t)k + lane);
                es.ids[base + p] = w;
                for (int q = 0; q < c.world; ++q)
                    if (q != c.rank) es.peer_ids[q][base + p] = w;
            }
            pos += __popc(bal);
        }
    }
    __syncthreads();            // the CTA's own segment is now readable by all its threads
}

// ---------------------------------------------------------------------------------------------
// View of one step's spike list straight from the segments: exclusive prefix sums of the
// segment counts in shared memory (cached per CTA under a tag, so the code objects of one phase
// build it once).  Global spike number g -> id by a binary search in shared memory.
// ---------------------------------------------------------------------------------------------
struct SpikeView {
    const unsigned long long* ids;   // ids of this slot
    const int32_t* seg_start;
    unsigned long long tag;  // id_tag of the step
    int remote_lo, remote_hi;   // segments [remote_lo, remote_hi) are the local rank's (no spinning)
    const int* pref;        // shared: pref[j - seg_lo] = number of spikes in segments [seg_lo, j)
    int seg_lo, nseg;       // segments covered
    int total;
    Control* ctrl;          // where a peer time-out is reported
};

__device__ int* view_storage(long long** tag) {
    __shared__ int s_pref[kMaxSegments + 1];
    __shared__ long long s_tag;
    *tag = &s_tag;
    return s_pref;
}
// Per-CTA memo for the state monitors of a launch: 1 = none of the recorded elements belongs to
// this CTA (the code object returns at once in every later step), 0 = some do, -1 = unknown.
constexpr int kMonitorMemo = 8;
__device__ __forceinline__ int* monitor_memo() {
    __shared__ int s_memo[kMonitorMemo];
    return s_memo;
}
// to be called once at kernel start
__device__ __forceinline__ void view_reset() {
    long long* tag;
    view_storage(&tag);
    if (threadIdx.x == 0) *tag = 0;
    if (threadIdx.x < kMonitorMemo) monitor_memo()[threadIdx.x] = -1;
    slice_cache_reset();
    __syncthreads();
}

__device__ __forceinline__ SpikeView view_build(const EventSpaceDev& es, int64_t timestep, const Ctx& c,
                                                bool local_only, Control* ctrl) {
    long long* tagp;
    int* pref = view_storage(&tagp);
    SpikeView v;
    const size_t slot = (size_t)ring_index(timestep, es.slots);
    v.ids = es.ids + slot * (size_t)es.N;
    const int32_t* cnt = es.cnt + slot * (size_t)es.nseg;
    v.seg_start = es.seg_start;
    v.tag = id_tag(timestep);
    v.remote_lo = c.rank * c.gnb;
    v.remote_hi = v.remote_lo + c.gnb;
    v.pref = pref;
    v.ctrl = ctrl;
    const bool local = local_only && c.world > 1;
    v.seg_lo = local ? c.rank * c.gnb : 0;
    v.nseg = local ? c.gnb : es.nseg;
    const long long tag = ((long long)(es.id * 2 + (local ? 1 : 0) + 1) << 44) ^ (timestep + 1);
    __syncthreads();
    if (*tagp != tag) {
        // block-wide exclusive scan of v.nseg counts (<= kMaxSegments); the count words of the
        // peers' segments are polled until they carry this step's tag (multi-GPU)
        const int want = step_tag(timestep);
        const bool poll = c.world > 1 && !local && timestep >= 0;
        __shared__ int s_wsum[kWarps];
        const int per = (v.nseg + kBlock - 1) / kBlock;
        const int first = (int)threadIdx.x * per;
        int vals[kMaxSegments / kBlock];
        int mine = 0;
#pragma unroll
        for (int k = 0; k < kMaxSegments / kBlock; ++k) {
            const int j = first + k;
            int w = 0;
            if (k < per && j < v.nseg) {
                const int* wp = cnt + v.seg_lo + j;
                if (poll && (v.seg_lo + j) / c.gnb != c.rank) {
                    // segment of a peer: spin until the word carries this step's tag
                    const long long t0 = clock64();
                    while (((w = ld_volatile_s32(wp)) >> 16) != want) {
                        if (clock64() - t0 > 40000000000LL || ld_volatile_s32(&ctrl->error)) {   // ~20 s
                            ctrl->error = 1;
                            raise_stop(ctrl);
                            w = 0;
                            break;
                        }
                    }
                    if (c.gbid == c.gnb - 1 && k == 0 && j == (c.rank == 0 ? c.gnb : 0)) {   // one sampled thread
                        ctrl->poll_cycles += (unsigned long long)(clock64() - t0);
                        ctrl->polls += 1;
                    }
                } else {
                    w = __ldcg(wp);
                    if (timestep < 0 || (w >> 16) != want) w = 0;   // slot holds no data of this step
                }
            }
            vals[k] = w & 0xffff;
            mine += vals[k];
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int x = s_wsum[w];
            if (w < warp) woff += x;
            tot += x;
        }
        int run = woff + incl - mine;
#pragma unroll
        for (int k = 0; k < kMaxSegments / kBlock; ++k) {
            const int j = first + k;
            if (k < per && j < v.nseg) { pref[j] = run; run += vals[k]; }
        }
        if (threadIdx.x == 0) { pref[v.nseg] = tot; *tagp = tag; }
        __syncthreads();
    }
    v.total = pref[v.nseg];
    return v;
}

// id stored at position `off` of segment `seg` (absolute segment index); a peer's word may
// still be in flight: spin until it carries the step's tag.  If it never does (~20 s: the peer
// died), the run is aborted through ctrl->error / the stop flag and the caller gets -1, an id no
// consumer accepts (pathways range-check the source, the monitors skip negative ids).
__device__ __forceinline__ int32_t view_load(const SpikeView& v, int seg, int off) {
    const unsigned long long* p = v.ids + v.seg_start[seg] + off;
    unsigned long long w = ld_volatile_u64(p);
    if (seg < v.remote_lo || seg >= v.remote_hi) {
        const long long t0 = clock64();
        while ((w & 0xffffffff00000000ULL) != v.tag) {
            if (clock64() - t0 > 40000000000LL || ld_volatile_s32(&v.ctrl->error)) {
                v.ctrl->error = 1;
                raise_stop(v.ctrl);
                return -1;
            }
            w = ld_volatile_u64(p);
        }
    }
    return (int32_t)(unsigned int)w;
}

// id of spike number g (0 <= g < total) of the view
__device__ __forceinline__ int32_t view_id(const SpikeView& v, int g) {
    int lo = 0, hi = v.nseg;           // last j with pref[j] <= g
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (v.pref[mid] <= g) lo = mid; else hi = mid;
    }
    return view_load(v, v.seg_lo + lo, g - v.pref[lo]);
}

// number of spikes of the view with id < x (for monitors of subgroups, spikemonitor.cpp:15-33)
__device__ __forceinline__ int view_count_below(const SpikeView& v, const EventSpaceDev& es, int x) {
    if (x <= v.seg_start[v.seg_lo]) return 0;
    if (x >= v.seg_start[v.seg_lo + v.nseg]) return v.total;
    int lo = 0, hi = v.nseg;           // segment with seg_start[lo] <= x < seg_start[lo+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (v.seg_start[v.seg_lo + mid] <= x) lo = mid; else hi = mid;
    }
    int a = 0, b = v.pref[lo + 1] - v.pref[lo];     // first position in the segment with id >= x
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (view_load(v, v.seg_lo + lo, mid) < x) a = mid + 1; else b = mid;
    }
    return v.pref[lo] + a;
}

// ---------------------------------------------------------------------------------------------
// Compaction: the reference layout of `_spikespace` (ids ascending in [0,count), count at [N])
// for step (timestep - lag), for the consumers that look back in time (delayed synapses) and
// for the host mirror.  CTA b copies the segments b, b+nb, ... (one per rank).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_segments(const EventSpaceDev& es, int64_t timestep, const Ctx& c,
                                                 Control* ctrl) {
    if (!es.need_compact) return;
    const int64_t s = timestep - es.lag;
    const SpikeView v = view_build(es, s, c, false, ctrl);
    int32_t* out = es.compact + (size_t)ring_index(s, es.slots) * (size_t)(es.N + 1);
    // one spike per thread: all loads of the step are in flight together (a CTA walking whole
    // segments would pay one round trip to L2 per segment)
    for (int g = c.bid * kBlock + (int)threadIdx.x; g < v.total; g += c.nb * kBlock) out[g] = view_id(v, g);
    if (c.bid == 0 && threadIdx.x == 0) out[es.N] = v.total;
}

// compact spike list of step `timestep` (valid once compact_segments ran for it)
__device__ __forceinline__ const int32_t* compact_slot(const EventSpaceDev& es, int64_t timestep) {
    return es.compact + (size_t)ring_index(timestep, es.slots) * (size_t)(es.N + 1);
}

// ---------------------------------------------------------------------------------------------
// atomics on arbitrary arithmetic types for the synaptic `x_post op= expr` statements
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(int32_t* p, int32_t v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(int64_t* p, int64_t v) {
    atomicAdd((unsigned long long*)p, (unsigned long long)v);
}
__device__ __forceinline__ void atomic_add(char* p, char v) {
    // bool/int8 accumulate: CAS on the containing word
    unsigned int* w = (unsigned int*)((size_t)p & ~(size_t)3);
    const unsigned int sh = ((size_t)p & 3) * 8;
    unsigned int old = *w, assumed;
    do {
        assumed = old;
        const unsigned int b = ((assumed >> sh) + (unsigned char)v) & 0xffu;
        old = atomicCAS(w, assumed, (assumed & ~(0xffu << sh)) | (b << sh));
    } while (old != assumed);
}
template <typename T, typename V>
__device__ __forceinline__ void atomic_add(T* p, V v) { atomic_add(p, (T)v); }

__device__ __forceinline__ void atomic_mul(double* p, double v) {
    unsigned long long* a = (unsigned long long*)p;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __double_as_longlong(__longlong_as_double(assumed) * v));
    } while (old != assumed);
}
__device__ __forceinline__ void atomic_mul(float* p, float v) {
    unsigned int* a = (unsigned int*)p;
    unsigned int old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __float_as_uint(__uint_as_float(assumed) * v));
    } while (old != assumed);
}
__device__ __forceinline__ void atomic_mul(int32_t* p, int32_t v) {
    int old = *p, assumed;
    do {
        assumed = old;
        old = atomicCAS(p, assumed, assumed * v);
    } while (old != assumed);
}
template <typename T, typename V>
__device__ __forceinline__ void atomic_mul(T* p, V v) { atomic_mul(p, (T)v); }
__device__ __forceinline__ void atomic_div(double* p, double v) {
    unsigned long long* a = (unsigned long long*)p;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __double_as_longlong(__longlong_as_double(assumed) / v));
    } while (old != assumed);
}
__device__ __forceinline__ void atomic_div(float* p, float v) {
    unsigned int* a = (unsigned int*)p;
    unsigned int old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __float_as_uint(__uint_as_float(assumed) / v));
    } while (old != assumed);
}
template <typename T, typename V>
__device__ __forceinline__ void atomic_div(T* p, V v) { atomic_div(p, (T)v); }

// ---------------------------------------------------------------------------------------------
// Counter-based RNG for in-loop rand()/randn() (Philox4x32-10).  The reference draws from one
// sequential mt19937 stream (objects.cpp:426-479), which no parallel device can reproduce; the
// reference itself disclaims cross-target reproducibility (docs_sphinx/advanced/random.rst).
// Stream key = (seed, code object id); counter = (element, timestep, call number).
// ---------------------------------------------------------------------------------------------
struct Rng {
    unsigned int key0, key1;
    unsigned int c_idx, c_step_lo, c_step_hi, c_call;
    double spare;
    int has_spare;
};

__device__ __forceinline__ void philox_round(unsigned int& c0, unsigned int& c1, unsigned int& c2,
                                             unsigned int& c3, unsigned int k0, unsigned int k1) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__device__ __forceinline__ void philox4x32(unsigned int c0, unsigned int c1, unsigned int c2,
                                           unsigned int c3, unsigned int k0, unsigned int k1,
                                           unsigned int out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ Rng rng_init(unsigned long long seed, unsigned int stream, int64_t idx,
                                        int64_t timestep) {
    Rng r;
    r.key0 = (unsigned int)seed ^ (stream * 0x9E3779B9u);
    r.key1 = (unsigned int)(seed >> 32) ^ stream;
    r.c_idx = (unsigned int)idx;
    r.c_step_lo = (unsigned int)timestep;
    r.c_step_hi = (unsigned int)((unsigned long long)timestep >> 32) ^ (unsigned int)((unsigned long long)idx >> 32);
    r.c_call = 0;
    r.has_spare = 0;
    r.spare = 0.0;
    return r;
}

// uniform in [0,1) with 53 random bits -- same construction as the reference's rand()
// (objects.cpp:448-453: (a>>5, b>>6) -> (a*2^26+b)/2^53)
__device__ __forceinline__ double rng_uniform(Rng& r) {
    unsigned int o[4];
    philox4x32(r.c_idx, r.c_step_lo, r.c_step_hi, r.c_call++, r.key0, r.key1, o);
    const double a = (double)(o[0] >> 5), b = (double)(o[1] >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

// polar Box-Muller with a cached second value, as objects.cpp:455-478
__device__ __forceinline__ double rng_normal(Rng& r) {
    if (r.has_spare) {
        r.has_spare = 0;
        return r.spare;
    }
    double x1, x2, r2;
    do {
        x1 = 2.0 * rng_uniform(r) - 1.0;
        x2 = 2.0 * rng_uniform(r) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = sqrt(-2.0 * log(r2) / r2);
    r.spare = f * x1;
    r.has_spare = 1;
    return f * x2;
}

// ---------------------------------------------------------------------------------------------
// poisson(lam) for in-loop code: the reference's two samplers (cpp_generator.py:661-751, the
// legacy numpy algorithms) on the element's Philox stream.  lam < 10: multiply uniforms until
// the product drops below exp(-lam); lam >= 10: Hoermann's transformed rejection with squeeze
// (PTRS), with the log-gamma of its acceptance test from the Stirling series after shifting the
// argument above 7.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double log_gamma_stirling(double x) {
    if (x == 1.0 || x == 2.0) return 0.0;
    const double c[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
                          -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
                          6.410256410256410e-03, -2.955065359477124e-02, 1.796443723688307e-01,
                          -1.39243221690590e+00};
    int shift = 0;
    double y = x;
    if (x <= 7.0) { shift = (int)(7.0 - x); y = x + shift; }
    const double inv2 = 1.0 / (y * y);
    double series = c[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) series = series * inv2 + c[k];
    double lg = series / y + 0.5 * log(6.283185307179586) + (y - 0.5) * log(y) - y;
    for (int k = 0; k < shift; ++k) { y -= 1.0; lg -= log(y); }
    return lg;
}

__device__ __forceinline__ int32_t rng_poisson(Rng& r, double lam) {
    if (lam == 0.0) return 0;
    if (lam < 10.0) {
        const double floor_ = exp(-lam);
        int32_t n = 0;
        double prod = rng_uniform(r);
        while (prod > floor_) { ++n; prod *= rng_uniform(r); }
        return n;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
    const double v_r = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        const double u = rng_uniform(r) - 0.5;
        const double v = rng_uniform(r);
        const double us = 0.5 - fabs(u);
        const int32_t k = (int32_t)floor((2.0 * a / us + b) * u + lam + 0.43);
        if (us >= 0.07 && v <= v_r) return k;
        if (k < 0 || (us < 0.013 && v > us)) continue;
        if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - log_gamma_stirling(k + 1.0))
            return k;
    }
}

}  // namespace b200
