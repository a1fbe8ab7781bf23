// b200_connect.cuh -- synapse creation on the device ("sharded construction").
//
// Reference behaviour restated here (brian-team/brian2):
//   brian2/devices/cpp_standalone/templates/synapses_create_generator.cpp:35-229
//       outer loop over the source (or target) neurons, inner iterator `range(low, high, step)`
//       or `sample(low, high, step, p=...)`; for p < 0.25 the sampler draws the gap to the next
//       accepted candidate from a geometric distribution (":155-173": jump = floor(log(u)/log(1-p))),
//       otherwise it tests every candidate with one uniform number.
//   brian2/synapses/synapses.py:2074-2275   (`_add_synapses_generator`: the four code blocks)
//
// Design (B200 first): the reference draws every random number of a connect() call from ONE
// sequential mt19937 stream, so synapse k cannot be known before synapses 0..k-1 -- the whole
// network has to be built by one thread and, on several GPUs, by every rank (16 GB of host
// arrays per rank at 10^9 synapses).  Here every ROW (outer index) owns a counter-based Philox
// stream keyed by (seed, connect call, row): rows are independent, one thread walks one row,
// and the result does not depend on the number of threads, CTAs or GPUs.  A rank keeps only the
// synapses whose postsynaptic neuron it owns, so the network is never materialised in one
// place.  Two passes over the same streams (count, prefix sum, fill) give the reference's
// (row ascending, candidate ascending) synapse order without atomics or sorting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <vector>

#include "b200_host.h"
#include "b200_runtime.cuh"

namespace b200 {

// arguments shared by both passes of a connect kernel
struct ConnectArgs {
    unsigned long long seed;
    unsigned int stream;        // one stream id per connect() call (crc of its name + call number)
    long long n_outer;          // rows
    long long post_lo, post_hi; // postsynaptic neurons (absolute index in the parent group) kept here
    int fill;                   // 0: count per row, 1: write the synapses
    const long long* row_start; // [n_outer] first output slot of every row (fill pass)
    int* row_count;             // [n_outer] (count pass)
    int32_t* pre;               // outputs of the fill pass
    int32_t* post;
    int* n_outgoing;            // [n_pre + source offset]   += per synapse (fill pass)
    int* n_incoming;            // [n_post + target offset]
    int* error;                 // != 0: an index left its valid range (1) / invalid sample size (2)
};

// `range(low, high, step)` / `sample(low, high, step, p)` of a connect() generator expression.
// next() hands out the candidates in ascending iteration order.
struct CandidateIter {
    long k, high, step;
    int sign;
    int mode;           // 0: every candidate, 1: Bernoulli test per candidate, 2: geometric jumps
    double p, inv_log1mp;
    bool done;

    __device__ __forceinline__ void init_range(long low, long high_, long step_) {
        k = low - step_; high = high_; step = step_;
        sign = step_ > 0 ? 1 : -1;
        mode = 0; p = 1.0; inv_log1mp = 0.0;
        done = step_ == 0;
    }
    __device__ __forceinline__ void init_sample(long low, long high_, long step_, double p_) {
        init_range(low, high_, step_);
        p = p_;
        if (p_ <= 0.0) { done = true; return; }
        // same switch-over as the reference (synapses_create_generator.cpp:157): jumps pay off
        // when fewer than a quarter of the candidates are accepted
        mode = p_ < 0.25 ? 2 : 1;
        inv_log1mp = 1.0 / log(1.0 - p_);
    }
    __device__ __forceinline__ bool next(long& out, Rng& rng) {
        while (!done) {
            k += step;
            if (sign * k >= sign * high) { done = true; break; }
            if (mode == 2) {
                const double u = rng_uniform(rng);
                if (u == 0.0) { done = true; break; }
                // number of rejected candidates in front of the next accepted one
                const double gap = floor(log(u) * inv_log1mp);
                if (gap >= 4.0e18) { done = true; break; }
                k += (long)gap * step;
                if (sign * k >= sign * high) { done = true; break; }
            } else if (mode == 1) {
                if (rng_uniform(rng) >= p) continue;
            }
            out = k;
            return true;
        }
        return false;
    }
};

// stands in for a group's index array `i` (an arange array) inside connect kernels
struct IdentityIndex {
    int start;
    __device__ __forceinline__ int32_t operator[](long long k) const { return (int32_t)(start + k); }
};

// what one row does with an accepted (pre, post, multiplicity) triple
struct RowSink {
    const ConnectArgs& a;
    long long slot;     // fill pass: next output slot of the row
    int count;          // count pass
    __device__ __forceinline__ RowSink(const ConnectArgs& a_, long long row)
        : a(a_), slot(a_.fill ? a_.row_start[row] : 0), count(0) {}
    __device__ __forceinline__ void emit(int32_t pre_idx, int32_t post_idx, int n) {
        if (post_idx < a.post_lo || post_idx >= a.post_hi) return;   // another rank's neuron
        if (!a.fill) { count += n; return; }
        for (int r = 0; r < n; ++r) {
            a.pre[slot] = pre_idx;
            a.post[slot] = post_idx;
            ++slot;
        }
        atomicAdd(a.n_outgoing + pre_idx, n);
        atomicAdd(a.n_incoming + post_idx, n);
    }
    __device__ __forceinline__ void finish(long long row) {
        if (!a.fill) a.row_count[row] = count;
    }
};

// ---------------------------------------------------------------------------------------------
// host driver: count pass, prefix sum, fill pass, append to the host mirrors
// ---------------------------------------------------------------------------------------------
struct ConnectResult {
    size_t created = 0;
    double seconds = 0.0;
};

// `launch(args)` starts the connect kernel of one code object on state().stream.
template <typename Launch>
inline ConnectResult connect_on_device(Launch launch, unsigned int stream_id, long long n_outer,
                                       int64_t n_post_parent, size_t n_pre_total, size_t n_post_total,
                                       std::vector<int32_t>& pre, std::vector<int32_t>& post,
                                       std::vector<int32_t>& n_incoming, std::vector<int32_t>& n_outgoing) {
    runtime_init();
    ensure_seed();
    RuntimeState& st = state();
    cudaEvent_t e0, e1;
    B200_CUDA(cudaEventCreate(&e0));
    B200_CUDA(cudaEventCreate(&e1));
    ConnectArgs a;
    memset(&a, 0, sizeof(a));
    a.seed = st.seed;
    a.stream = stream_id;
    a.n_outer = n_outer;
    a.post_lo = 0;
    a.post_hi = INT64_MAX;
    if (st.world > 1) {
        int64_t lo, hi;
        EventSpace::rank_range_host(n_post_parent, st.rank, st.world, lo, hi);
        a.post_lo = lo;
        a.post_hi = hi;
    }
    n_incoming.resize(n_post_total);
    n_outgoing.resize(n_pre_total);
    const size_t rows = (size_t)std::max<long long>(n_outer, 1);
    int* d_count = (int*)dev_alloc(rows * sizeof(int));
    long long* d_start = (long long*)dev_alloc(rows * sizeof(long long));
    int* d_error = (int*)dev_alloc(sizeof(int));
    B200_CUDA(cudaMemset(d_count, 0, rows * sizeof(int)));
    B200_CUDA(cudaMemset(d_error, 0, sizeof(int)));
    a.row_count = d_count;
    a.row_start = d_start;
    a.error = d_error;
    B200_CUDA(cudaEventRecord(e0, st.stream));
    a.fill = 0;
    launch(a);
    B200_CUDA(cudaGetLastError());
    std::vector<int> count(rows);
    B200_CUDA(cudaMemcpyAsync(count.data(), d_count, rows * sizeof(int), cudaMemcpyDeviceToHost, st.stream));
    int err = 0;
    B200_CUDA(cudaMemcpyAsync(&err, d_error, sizeof(int), cudaMemcpyDeviceToHost, st.stream));
    B200_CUDA(cudaStreamSynchronize(st.stream));
    auto cleanup = [&]() { dev_free(d_count); dev_free(d_start); dev_free(d_error); };
    if (err) {
        cleanup();
        throw std::runtime_error(err == 1 ? "b200 connect: tried to create a synapse to/from a neuron outside the valid index range"
                                          : "b200 connect: invalid sample size");
    }
    std::vector<long long> start(rows);
    long long total = 0;
    for (size_t r = 0; r < rows; ++r) { start[r] = total; total += count[r]; }
    if ((size_t)total + pre.size() >= (size_t)INT32_MAX) {
        cleanup();
        throw std::runtime_error("b200 connect: more than 2^31-1 synapses in one Synapses object on one GPU");
    }
    ConnectResult res;
    res.created = (size_t)total;
    int32_t* d_pre = (int32_t*)dev_alloc(std::max<size_t>(1, (size_t)total) * sizeof(int32_t));
    int32_t* d_post = (int32_t*)dev_alloc(std::max<size_t>(1, (size_t)total) * sizeof(int32_t));
    int* d_nout = (int*)dev_alloc(std::max<size_t>(1, n_pre_total) * sizeof(int));
    int* d_nin = (int*)dev_alloc(std::max<size_t>(1, n_post_total) * sizeof(int));
    B200_CUDA(cudaMemcpyAsync(d_start, start.data(), rows * sizeof(long long), cudaMemcpyHostToDevice, st.stream));
    if (n_pre_total) B200_CUDA(cudaMemcpyAsync(d_nout, n_outgoing.data(), n_pre_total * sizeof(int), cudaMemcpyHostToDevice, st.stream));
    if (n_post_total) B200_CUDA(cudaMemcpyAsync(d_nin, n_incoming.data(), n_post_total * sizeof(int), cudaMemcpyHostToDevice, st.stream));
    a.fill = 1;
    a.pre = d_pre;
    a.post = d_post;
    a.n_outgoing = d_nout;
    a.n_incoming = d_nin;
    launch(a);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(e1, st.stream));
    const size_t old = pre.size();
    pre.resize(old + (size_t)total);
    post.resize(old + (size_t)total);
    if (total) {
        B200_CUDA(cudaMemcpyAsync(pre.data() + old, d_pre, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, st.stream));
        B200_CUDA(cudaMemcpyAsync(post.data() + old, d_post, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, st.stream));
    }
    if (n_pre_total) B200_CUDA(cudaMemcpyAsync(n_outgoing.data(), d_nout, n_pre_total * sizeof(int), cudaMemcpyDeviceToHost, st.stream));
    if (n_post_total) B200_CUDA(cudaMemcpyAsync(n_incoming.data(), d_nin, n_post_total * sizeof(int), cudaMemcpyDeviceToHost, st.stream));
    B200_CUDA(cudaStreamSynchronize(st.stream));
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    res.seconds = 1e-3 * ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(d_pre); dev_free(d_post); dev_free(d_nout); dev_free(d_nin);
    cleanup();
    st.connect_launches += 2;
    st.connect_seconds += res.seconds;
    st.connect_synapses += (double)total;
    return res;
}

}  // namespace b200
