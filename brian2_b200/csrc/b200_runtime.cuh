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
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// Grid barrier for the persistent kernel.  Monotonic 64-bit arrival counter (never reset inside
// a launch); `target` is thread-0 private state.  Thread 0 fences on both sides (the gpu-scope
// fence also invalidates this SM's L1, so plain loads after the barrier see other CTAs' writes).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long& target,
                                             int nb) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += (unsigned long long)nb;
        __threadfence();
        atomicAdd(counter, 1ULL);
        while (ld_acquire_u64(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Owned-slice partition
// ---------------------------------------------------------------------------------------------
struct Slice {
    int64_t lo;   // first element of this warp's slice
    int64_t hi;   // lo + slice length (may exceed N: guard with idx < N)
};

__device__ __forceinline__ int64_t owned_chunk(int64_t N, int nb) {
    int64_t per = (N + nb - 1) / nb;
    return (per + (kBlock - 1)) & ~(int64_t)(kBlock - 1);
}

__device__ __forceinline__ Slice owned_slice(int64_t N, const Ctx& c) {
    const int64_t chunk = owned_chunk(N, c.nb);
    const int64_t wchunk = chunk / kWarps;
    Slice s;
    s.lo = (int64_t)c.bid * chunk + (int64_t)(threadIdx.x >> 5) * wchunk;
    s.hi = s.lo + wchunk;
    return s;
}

// for (idx in my lane's elements) -- idx is guarded against N by the caller
#define B200_FOR_OWNED(IDX, N_, CTX)                                                        \
    const b200::Slice _b200_sl = b200::owned_slice((int64_t)(N_), (CTX));                    \
    for (int64_t IDX = _b200_sl.lo + (threadIdx.x & 31);                                    \
         IDX < _b200_sl.hi && IDX < (int64_t)(N_); IDX += 32)

// ---------------------------------------------------------------------------------------------
// Ordered compaction of owned-slice flags into an event space (ids ascending, count at [N]).
//   mask   bit k of lane l  <=>  element (slice.lo + 32*k + l) fired
//   ws     look-back workspace, one u64 per CTA: (epoch << 32) | count ; epoch must be unique
//          per call for this workspace and never 0.
// Called by ALL threads of ALL CTAs (CTAs without elements publish 0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_owned(unsigned long long mask, int niter, int64_t N,
                                              const Ctx& c, int32_t* __restrict__ eventspace,
                                              unsigned long long* ws, unsigned int epoch) {
    __shared__ int s_warp[kWarps];
    __shared__ int s_pred[kWarps];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const Slice sl = owned_slice(N, c);

    // Only CTAs that own elements take part (for small groups most of the grid has nothing to
    // compact and must not pay the look-back latency); the last owning CTA writes the count.
    const int64_t chunk = owned_chunk(N, c.nb);
    const int nact = (int)((N + chunk - 1) / chunk);
    if (c.bid >= nact) return;

    // 1. per-warp totals
    int mine = __popcll(mask);
    int wtotal = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wtotal += __shfl_xor_sync(0xffffffffu, wtotal, o);
    if (lane == 0) s_warp[warp] = wtotal;
    __syncthreads();

    // 2. CTA total -> publish ; exclusive warp offsets
    int woff = 0, ctotal = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int v = s_warp[w];
        if (w < warp) woff += v;
        ctotal += v;
    }
    if (threadIdx.x == 0 && c.bid + 1 < nact) {
        st_release_u64(&ws[c.bid], ((unsigned long long)epoch << 32) | (unsigned int)ctotal);
    }

    // 3. decoupled look-back: thread j < bid polls predecessor j (nb <= kBlock)
    int pred = 0;
    if ((int)threadIdx.x < c.bid) {
        unsigned long long v;
        do {
            v = ld_acquire_u64(&ws[threadIdx.x]);
        } while ((unsigned int)(v >> 32) != epoch);
        pred = (int)(unsigned int)(v & 0xffffffffu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pred += __shfl_xor_sync(0xffffffffu, pred, o);
    if (lane == 0) s_pred[warp] = pred;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) b += s_pred[w];
        s_base = b;
        if (c.bid == nact - 1) eventspace[N] = b + ctotal;
    }
    __syncthreads();

    // 4. write ids in ascending order: iteration-major, lane-minor inside the warp slice
    int pos = s_base + woff;
    if (wtotal > 0) {
        for (int k = 0; k < niter; ++k) {
            const bool f = (mask >> k) & 1ULL;
            const unsigned int bal = __ballot_sync(0xffffffffu, f);
            if (f) {
                const int p = pos + __popc(bal & ((1u << lane) - 1u));
                eventspace[p] = (int32_t)(sl.lo + 32 * (int64_t)k + lane);
            }
            pos += __popc(bal);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sorted-range helpers on an event space (ids ascending): first index with id >= x
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}


// spike list that a delay-bin has to deliver at `timestep`
__device__ __forceinline__ int ring_index(int64_t timestep, int slots) {
    int64_t s = timestep % slots;
    if (s < 0) s += slots;
    return (int)s;
}
__device__ __forceinline__ const int32_t* ring_slot(const int32_t* ring, int slots, int stride,
                                                    int64_t timestep) {
    int64_t s = timestep % slots;
    if (s < 0) s += slots;
    return ring + s * (int64_t)stride;
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


}  // namespace b200
