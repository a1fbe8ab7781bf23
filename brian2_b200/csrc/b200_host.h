// b200_host.h -- host-side runtime of the Brian2 `b200` device: device memory management, the
// device-resident spike ring / delay-binned CSR that replaces CSpikeQueue, monitor buffers.
//
// Reference behaviour restated here (brian-team/brian2):
//   brian2/synapses/spikequeue.h:48-105    CSpikeQueue::prepare  (delay rounding, per-source lists)
//   brian2/synapses/spikequeue.h:151-205   push / peek / advance  (order of delivery)
//   brian2/devices/cpp_standalone/templates/synapses_classes.cpp:15-88   SynapticPathway
//   brian2/devices/cpp_standalone/brianlib/dynamic_array.h            growth of monitor storage
#pragma once
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200_types.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// error handling: never exit() from the library; exceptions are caught at the C-ABI boundary
// ---------------------------------------------------------------------------------------------
inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %s (%d) at %s:%d: %s", cudaGetErrorString(e), (int)e,
                 file, line, what);
        throw std::runtime_error(buf);
    }
}
#define B200_CUDA(x) ::b200::cuda_check((x), #x, __FILE__, __LINE__)

struct RuntimeState {
    int device = 0;
    int rank = 0;
    int world = 1;
    int num_sms = 0;
    int grid = 0;                  // CTAs used by every generated kernel (co-resident)
    bool initialised = false;
    cudaStream_t stream = nullptr;
    Control* control = nullptr;        // device
    Control* control_host = nullptr;   // pinned
    volatile int* stop_request = nullptr;   // host-mapped flag polled by the persistent kernel
    int* stop_request_dev = nullptr;
    size_t bytes_allocated = 0;
    unsigned long long seed = 0;
    bool seeded = false;
    // multi-GPU plumbing: set through b200_set_comm (include/brian2_b200.h) before b200_run_main
    int (*allgather)(const void* send, void* recv, size_t nbytes_per_rank) = nullptr;
    unsigned long long launches = 0;   // kernels launched inside run loops
    double poll_cycles = 0, fence_cycles = 0, polls = 0;   // multi-GPU wait diagnostics
    double upload_seconds = 0.0, download_seconds = 0.0;
    size_t h2d_bytes = 0, d2h_bytes = 0;
    // synapse creation on the device (b200_connect.cuh)
    double connect_seconds = 0.0, connect_synapses = 0.0;
    unsigned long long connect_launches = 0;
    double prepare_seconds = 0.0;       // host time spent building delay-binned CSRs
    unsigned long long host_epoch = 1;  // bumped by everything that may change a host array
    // every pathway of the project delivers at least one step after the spike (decided at
    // upload): the step kernels of the 'd1' variant run (no end-of-step barrier, see device.py)
    bool all_delayed = false;
};

inline RuntimeState& state() {
    static RuntimeState s;
    return s;
}

inline void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    B200_CUDA(cudaMalloc(&p, bytes));
    state().bytes_allocated += bytes;
    return p;
}
inline void dev_free(void* p) {
    if (p) cudaFree(p);
}

inline void runtime_init() {
    RuntimeState& s = state();
    if (s.initialised) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error(
            "b200 device: no CUDA device available (this device has no CPU fallback)");
    // one process per GPU: rank/world come from b200_set_comm; the device from LOCAL_RANK
    const char* lr = getenv("LOCAL_RANK");
    s.device = lr ? atoi(lr) % ndev : 0;
    if (s.world > kMaxRanks) throw std::runtime_error("b200: at most 8 ranks (GPUs of one box)");
    if (s.world > 1 && !s.allgather)
        throw std::runtime_error("b200: world > 1 needs an allgather callback (b200_set_comm)");
    B200_CUDA(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, s.device));
    s.num_sms = prop.multiProcessorCount;
    B200_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    s.control = (Control*)dev_alloc(sizeof(Control));
    B200_CUDA(cudaMemset(s.control, 0, sizeof(Control)));
    B200_CUDA(cudaHostAlloc((void**)&s.control_host, sizeof(Control), cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc((void**)&s.stop_request, sizeof(int), cudaHostAllocMapped));
    *s.stop_request = 0;
    B200_CUDA(cudaHostGetDevicePointer((void**)&s.stop_request_dev, (void*)s.stop_request, 0));
    s.initialised = true;
}

// ---------------------------------------------------------------------------------------------
// plain arrays: host mirror <-> device
// ---------------------------------------------------------------------------------------------
template <typename T>
inline void upload_array(T*& dev, const T* host, size_t n) {
    if (!dev) dev = (T*)dev_alloc(n * sizeof(T));
    if (n) B200_CUDA(cudaMemcpy(dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
    state().h2d_bytes += n * sizeof(T);
}
template <typename T>
inline void download_array(T* host, const T* dev, size_t n) {
    if (n && dev) B200_CUDA(cudaMemcpy(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost));
    state().d2h_bytes += n * sizeof(T);
}

// dynamic 1-d arrays (std::vector on the host).  `cap` is the device capacity in elements.
template <typename T>
inline void upload_vector(T*& dev, size_t& cap, size_t& n, const std::vector<T>& host,
                          size_t min_cap = 0) {
    const size_t need = std::max(host.size(), min_cap);
    if (!dev || cap < need) {
        dev_free(dev);
        cap = std::max<size_t>(need, 16);
        dev = (T*)dev_alloc(cap * sizeof(T));
    }
    n = host.size();
    if (n) B200_CUDA(cudaMemcpy(dev, host.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    state().h2d_bytes += n * sizeof(T);
}
template <typename T>
inline void download_vector(std::vector<T>& host, const T* dev, size_t n) {
    host.resize(n);
    if (n && dev) B200_CUDA(cudaMemcpy(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
    state().d2h_bytes += n * sizeof(T);
}
// Monitor records (append-only, read-only for the user): host and device copies agree on the
// first `synced` elements after every download, so a later run() neither sends them to the
// device again nor fetches them a second time -- only the records of the run that just ended
// cross PCIe.
template <typename T>
inline void upload_records(T*& dev, size_t& cap, size_t& n, const std::vector<T>& host, size_t min_cap,
                           size_t& synced) {
    if (dev && host.size() == synced && n == synced && cap >= std::max(host.size(), min_cap)) return;
    upload_vector(dev, cap, n, host, min_cap);
    synced = host.size();
}
template <typename T>
inline void download_records(std::vector<T>& host, const T* dev, size_t n, size_t& synced) {
    const size_t from = (host.size() == synced && n >= synced) ? synced : 0;
    host.resize(n);
    if (n > from && dev)
        B200_CUDA(cudaMemcpy(host.data() + from, dev + from, (n - from) * sizeof(T), cudaMemcpyDeviceToHost));
    state().d2h_bytes += (n - from) * sizeof(T);
    synced = n;
}
// grow a device append buffer, preserving the first `used` elements
template <typename T>
inline void grow_buffer(T*& dev, size_t& cap, size_t used, size_t new_cap) {
    if (new_cap <= cap && dev) return;
    T* nd = (T*)dev_alloc(new_cap * sizeof(T));
    if (dev && used)
        B200_CUDA(cudaMemcpy(nd, dev, used * sizeof(T), cudaMemcpyDeviceToDevice));
    dev_free(dev);
    dev = nd;
    cap = new_cap;
}

// ---------------------------------------------------------------------------------------------
// host-side collectives over the callback (setup / teardown only, never inside the step loop)
// ---------------------------------------------------------------------------------------------
inline void host_allgather(const void* send, void* recv, size_t nbytes) {
    RuntimeState& s = state();
    if (s.world <= 1) { memcpy(recv, send, nbytes); return; }
    if (s.allgather(send, recv, nbytes) != 0) throw std::runtime_error("b200: allgather callback failed");
}
// Seed of the in-loop Philox streams when the script never called seed(): drawn from the
// operating system like the reference's generators (objects.cpp:437-440 seeds from
// std::random_device), so that unseeded runs see different noise; on several GPUs every rank
// uses rank 0's draw (the shards of one network must agree on the streams).
inline void ensure_seed() {
    RuntimeState& s = state();
    if (s.seeded) return;
    std::random_device rd;
    unsigned long long mine = ((unsigned long long)rd() << 32) ^ (unsigned long long)rd();
    unsigned long long all[kMaxRanks] = {0};
    host_allgather(&mine, all, sizeof(mine));
    s.seed = all[0];
    s.seeded = true;
}
inline void host_barrier() {
    if (state().world <= 1) return;
    char one = 1, all[kMaxRanks];
    host_allgather(&one, all, 1);
}

// ---------------------------------------------------------------------------------------------
// Event space: the last `slots` spike lists of one group, as per-CTA segments and (a few steps
// later) as the reference's compact `_spikespace` layout -- see EventSpaceDev in b200_types.h.
// ---------------------------------------------------------------------------------------------
struct EventSpace {
    unsigned long long* ids = nullptr;   // [slots][N]     tagged ids, per-CTA segments
    int32_t* cnt = nullptr;              // [slots][nseg]  tagged segment counts
    int32_t* compact = nullptr;          // [slots][N+1]   reference layout
    int32_t* seg_start = nullptr;
    std::vector<int32_t> seg_start_host;
    int slots = 0, N = 0, nb = 0, nseg = 0, id = 0;
    int max_delay = 0;         // over all pathways reading this event space
    int min_delay = 1 << 30;   // ditto (1<<30: no pathway)
    void* peer_ids[kMaxRanks] = {0};
    void* peer_cnt[kMaxRanks] = {0};
    bool peers_open = false;
    bool compact_always = false;   // a consumer walks the compact list of the current step

    void require(int dmin, int dmax) {
        max_delay = std::max(max_delay, dmax);
        min_delay = std::min(min_delay, dmin);
    }
    // On several GPUs a step is compacted only when a consumer can first need it (min delay - 1
    // steps later), so that the wait for the peers' segments never stalls the step loop.
    // 'd1' variant (all delays >= 1 step): the list of the previous step is read from the
    // segments, compacted lists are first needed two steps after the spike.
    int lag() const {
        if (min_delay == (1 << 30)) return 0;
        if (state().all_delayed) return std::max(0, min_delay - 2);
        if (state().world <= 1) return 0;
        return std::max(0, min_delay - 1);
    }
    int required_slots() const { return 2 * (max_delay + 1) + 2; }

    void close_peers() {
        for (int q = 0; q < kMaxRanks; ++q) {
            if (peer_ids[q] && q != state().rank) cudaIpcCloseMemHandle(peer_ids[q]);
            if (peer_cnt[q] && q != state().rank) cudaIpcCloseMemHandle(peer_cnt[q]);
            peer_ids[q] = peer_cnt[q] = nullptr;
        }
        peers_open = false;
    }

    template <typename T>
    static T* realloc_ring(T* old, int old_slots, size_t old_stride, int new_slots, size_t new_stride,
                           int64_t timestep, bool keep) {
        T* nd = (T*)dev_alloc((size_t)new_slots * new_stride * sizeof(T));
        B200_CUDA(cudaMemset(nd, 0, (size_t)new_slots * new_stride * sizeof(T)));
        if (old && keep) {
            // same layout: carry the history of the last steps over (slots only ever grow)
            for (int back = 1; back <= old_slots && back < new_slots; ++back) {
                const int64_t s = timestep - back;
                int64_t os = s % old_slots, ns = s % new_slots;
                if (os < 0) os += old_slots;
                if (ns < 0) ns += new_slots;
                B200_CUDA(cudaMemcpy(nd + ns * new_stride, old + os * old_stride, old_stride * sizeof(T),
                                     cudaMemcpyDeviceToDevice));
            }
        }
        dev_free(old);
        return nd;
    }

    // (re)allocate; keeps the history of the last steps before `timestep`
    void ensure(int N_, int nb_, int64_t timestep) {
        RuntimeState& st = state();
        const int need = required_slots();
        if (ids && slots >= need && N == N_ && nb == nb_) return;
        const int new_nseg = st.world * nb_;
        {   // the thresholder keeps one bit per owned element of a lane in a 64-bit mask
            int64_t lo, hi;
            rank_range_host(N_, st.rank, st.world, lo, hi);
            const int64_t tasks = (hi - lo + 31) >> 5, warps = (int64_t)nb_ * kWarps;
            if ((tasks + warps - 1) / warps > kMaxOwnedIters)
                throw std::runtime_error("b200: group too large for one GPU (more than 64 x 32 elements per "
                                         "warp of the grid); partition the network over more GPUs");
        }
        const bool keep = ids && N == N_ && nb == nb_;
        close_peers();
        ids = realloc_ring(ids, slots, (size_t)N, need, (size_t)N_, timestep, keep);
        cnt = realloc_ring(cnt, slots, (size_t)nseg, need, (size_t)new_nseg, timestep, keep);
        compact = realloc_ring(compact, slots, (size_t)N + 1, need, (size_t)N_ + 1, timestep, keep);
        slots = need; N = N_; nb = nb_; nseg = new_nseg;
        // first neuron of every segment (same arithmetic as owned_cta on the device)
        std::vector<int32_t> start(nseg + 1);
        for (int q = 0; q < st.world; ++q) {
            int64_t lo, hi;
            rank_range_host(N, q, st.world, lo, hi);
            for (int b = 0; b < nb; ++b)
                start[q * nb + b] = (int32_t)warp_first_host(lo, hi, (int64_t)b * kWarps, (int64_t)nb * kWarps);
        }
        start[nseg] = N;
        seg_start_host = start;
        dev_free(seg_start);
        seg_start = (int32_t*)dev_alloc((nseg + 1) * sizeof(int32_t));
        B200_CUDA(cudaMemcpy(seg_start, start.data(), (nseg + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    }

    static void rank_range_host(int64_t N, int rank, int world, int64_t& lo, int64_t& hi) {
        int64_t per = (N + world - 1) / world;
        per = (per + 31) & ~(int64_t)31;
        lo = (int64_t)rank * per; if (lo > N) lo = N;
        hi = lo + per; if (hi > N) hi = N;
    }
    static int64_t warp_first_host(int64_t lo, int64_t hi, int64_t g, int64_t G) {
        const int64_t T = (hi - lo + 31) >> 5;
        int64_t e = lo + 32 * ((g * T) / G);
        return e < hi ? e : hi;
    }

    // multi-GPU: map every peer's ring (CUDA IPC, one handle exchange per allocation)
    void open_peers() {
        RuntimeState& st = state();
        if (st.world <= 1 || peers_open) return;
        B200_CUDA(cudaDeviceSynchronize());
        struct Handles { cudaIpcMemHandle_t ids, cnt; } mine, all[kMaxRanks];
        B200_CUDA(cudaIpcGetMemHandle(&mine.ids, ids));
        B200_CUDA(cudaIpcGetMemHandle(&mine.cnt, cnt));
        host_allgather(&mine, all, sizeof(Handles));
        for (int q = 0; q < st.world; ++q) {
            if (q == st.rank) { peer_ids[q] = ids; peer_cnt[q] = cnt; continue; }
            B200_CUDA(cudaIpcOpenMemHandle(&peer_ids[q], all[q].ids, cudaIpcMemLazyEnablePeerAccess));
            B200_CUDA(cudaIpcOpenMemHandle(&peer_cnt[q], all[q].cnt, cudaIpcMemLazyEnablePeerAccess));
        }
        peers_open = true;
    }

    EventSpaceDev view() const {
        RuntimeState& st = state();
        EventSpaceDev v;
        memset(&v, 0, sizeof(v));
        v.ids = ids; v.cnt = cnt; v.compact = compact; v.seg_start = seg_start;
        v.slots = slots; v.N = N; v.nseg = nseg; v.lag = lag(); v.id = id;
        // the reference layout is only built when somebody reads it (delayed or serial pathways)
        v.need_compact = (compact_always || max_delay > (state().all_delayed ? 1 : 0)) ? 1 : 0;
        int64_t lo, hi;
        rank_range_host(N, st.rank, st.world, lo, hi);
        v.rank_lo = (int)lo; v.rank_hi = (int)hi;
        for (int q = 0; q < st.world; ++q) {
            v.peer_ids[q] = (unsigned long long*)peer_ids[q];
            v.peer_cnt[q] = (int32_t*)peer_cnt[q];
        }
        return v;
    }
    // Host mirror of the list of step `timestep` in the reference layout (ids ascending in
    // [0, count), count at [N]), assembled from the segments: works whether or not the device
    // ever built the compact form (it only does when a delayed / serial pathway reads it).
    void download_step(int32_t* host, int64_t timestep) const {
        if (!ids || N <= 0) return;
        int64_t s = timestep % slots;
        if (s < 0) s += slots;
        std::vector<int32_t> c(nseg);
        std::vector<unsigned long long> w((size_t)N);
        B200_CUDA(cudaMemcpy(c.data(), cnt + s * (size_t)nseg, nseg * sizeof(int32_t), cudaMemcpyDeviceToHost));
        B200_CUDA(cudaMemcpy(w.data(), ids + s * (size_t)N, (size_t)N * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        int64_t m = (timestep + 1) % 32767;          // == b200::step_tag (b200_runtime.cuh)
        if (m < 0) m += 32767;
        const int want = (int)m + 1;
        int n = 0;
        for (int j = 0; j < nseg; ++j) {
            if (timestep < 0 || (c[j] >> 16) != want) continue;
            const int k = c[j] & 0xffff;
            for (int q = 0; q < k && n < N; ++q) host[n++] = (int32_t)(unsigned int)w[(size_t)seg_start_host[j] + q];
        }
        host[N] = n;
        state().d2h_bytes += nseg * sizeof(int32_t) + (size_t)N * sizeof(unsigned long long);
    }
    const int32_t* compact_slot_ptr(int64_t timestep) const {
        int64_t s = timestep % slots;
        if (s < 0) s += slots;
        return compact + s * ((size_t)N + 1);
    }
};

// ---------------------------------------------------------------------------------------------
// Pathway: device-resident replacement of SynapticPathway + CSpikeQueue.
// ---------------------------------------------------------------------------------------------
class Pathway {
public:
    std::vector<int>& sources;     // same constructor contract as the reference class
    int spikes_start, spikes_stop;
    int Nsource = 0, Ntarget = 0;
    int max_delay = 0;
    int nbins = 0;
    size_t n_synapses = 0;
    bool identity = true;
    bool prepared = false;
    unsigned long long built_epoch = 0;
    double built_dt = 0.0;
    std::vector<int> bin_delay;
    // device storage
    int* d_bin_delay = nullptr;
    int* d_rowptr = nullptr;
    int* d_syn_ids = nullptr;
    int* d_csr_target = nullptr;
    unsigned long long* d_events = nullptr;
    unsigned int* d_tickets = nullptr;
    int* d_hits = nullptr;
    int hits_n = 0;
    int* d_tileptr = nullptr;      // b200_tiles.cuh
    int tile_stride = 0, tiles_grid = 0;
    int tile_n = 0;                // size of the target group of a countable pathway (0: not countable)
    int counted_n = 0;             // > 0: counted pathway (hits + apply pass)
    bool forward = false;          // counted + every delay >= 1 step: "forward" layout, see prepare()
    int hits_alloc_slots = 0;
    int hits_slots = 2;            // rings of the hit counters: 2 (by step parity) or max_delay + 1
    static bool& allow_forward() { static bool v = true; return v; }
    bool tiles_tried = false;      // since the CSR was last built
    EventSpace* es = nullptr;
    size_t n_owned = 0;            // synapses stored on this rank (post neuron owned)

    // `counted_n` > 0: the synaptic code of this pathway is "counted" (events per target element,
    // applied by the owner; size of the target group) -- known when the project is generated
    Pathway(std::vector<int>& _sources, int _spikes_start, int _spikes_stop, int counted_n = 0)
        : sources(_sources), spikes_start(_spikes_start), spikes_stop(_spikes_stop), counted_n(counted_n) {}

    ~Pathway() { release(); }

    void release() {
        dev_free(d_bin_delay); dev_free(d_rowptr); dev_free(d_syn_ids); dev_free(d_csr_target);
        dev_free(d_tileptr);
        d_bin_delay = d_rowptr = d_syn_ids = d_csr_target = d_tileptr = nullptr;
        tiles_tried = false;
    }

    // Forward layout (see prepare()): synapses ordered by (source, delay bin, synapse index);
    // rowptr[s * (nbins + 1) + b] = first slot of bin b of source s, rowptr[s * (nbins + 1) + nbins]
    // = one past the last slot of source s; csr_target[slot] = bin << 27 | target.  Pure host
    // code (tests/cuda/forward_csr_test.cpp checks it on the CPU).  `bins` = nullptr: one bin.
    template <typename Owned>
    static size_t build_forward_csr(int nsrc, int nbins, int spikes_start, const int* srcs, const int* targets,
                                    const int* bins, size_t n_syn, Owned owned, std::vector<int>& rowptr,
                                    std::vector<int>& csr_target) {
        const size_t stride = (size_t)nbins + 1;
        rowptr.assign((size_t)nsrc * stride + 1, 0);
        std::vector<int> count((size_t)nsrc * nbins + 1, 0);
        size_t kept = 0;
        for (size_t i = 0; i < n_syn; ++i) {
            const int sidx = srcs[i] - spikes_start;
            if (sidx < 0 || sidx >= nsrc) throw std::runtime_error("b200: synapse source outside pathway source range");
            if (!owned(i)) continue;
            count[(size_t)sidx * nbins + (bins ? bins[i] : 0)]++;
            kept++;
        }
        int run = 0;
        for (int sidx = 0; sidx < nsrc; ++sidx) {
            for (int b = 0; b < nbins; ++b) {
                rowptr[(size_t)sidx * stride + b] = run;
                run += count[(size_t)sidx * nbins + b];
            }
            rowptr[(size_t)sidx * stride + nbins] = run;
        }
        csr_target.assign(kept, 0);
        std::vector<int> cursor((size_t)nsrc * nbins);
        for (int sidx = 0; sidx < nsrc; ++sidx)
            for (int b = 0; b < nbins; ++b) cursor[(size_t)sidx * nbins + b] = rowptr[(size_t)sidx * stride + b];
        for (size_t i = 0; i < n_syn; ++i) {
            if (!owned(i)) continue;
            const int sidx = srcs[i] - spikes_start, b = bins ? bins[i] : 0;
            csr_target[cursor[(size_t)sidx * nbins + b]++] = (int)(((unsigned int)b << 27) | (unsigned int)targets[i]);
        }
        return kept;
    }

    // Build the delay-binned CSR.  Delay rounding as CSpikeQueue::prepare (spikequeue.h:90):
    // steps = (int)(delay/dt + 0.5); n_delays == 1 means one delay for all synapses (:104).
    // Slot order = (delay bin asc, source asc, synapse index asc); walking the bins from the
    // largest delay to the smallest reproduces the reference's delivery order within a step
    // (entries pushed earlier sit first in a bucket, spikequeue.h:157-190).
    // `post_is_source`: the pathway listens to the POSTsynaptic group (on_post), i.e. `srcs` are
    // the postsynaptic ends.  On several GPUs a rank stores only the synapses whose postsynaptic
    // neuron it owns ([post_lo, post_hi) of the postsynaptic group's parent).
    template <typename scalar>
    void prepare(int n_source, int n_target, const scalar* real_delays, size_t n_delays,
                 const int* srcs, const int* targets, size_t n_syn, double dt, EventSpace* es_,
                 bool post_is_source, int64_t n_post_parent) {
        runtime_init();
        // Nothing on the host changed since this CSR was built (the generated main() bumps
        // host_epoch after every host-side write): keep it, the next run() reuses it as is.
        if (prepared && built_epoch == state().host_epoch && n_synapses == n_syn && es == es_ && built_dt == dt) {
            B200_CUDA(cudaMemset(d_tickets, 0, 2 * sizeof(unsigned int)));
            if (es) es->require(bin_delay.empty() ? 0 : bin_delay.front(), max_delay);
            return;
        }
        const auto _t0 = std::chrono::high_resolution_clock::now();
        release();
        built_epoch = state().host_epoch;
        built_dt = dt;
        Nsource = n_source;
        Ntarget = n_target;
        n_synapses = n_syn;
        es = es_;
        int64_t post_lo = 0, post_hi = INT64_MAX;
        if (state().world > 1)
            EventSpace::rank_range_host(n_post_parent, state().rank, state().world, post_lo, post_hi);
        const int* post_end = post_is_source ? srcs : targets;
        auto owned = [&](size_t i) -> bool {
            return state().world <= 1 || !post_end || (post_end[i] >= post_lo && post_end[i] < post_hi);
        };
        if (n_syn >= (size_t)INT32_MAX)
            throw std::runtime_error("b200: more than 2^31-1 synapses in one pathway shard");
        const int nsrc = spikes_stop - spikes_start;
        std::vector<int> dsteps;
        int scalar_delay = 0;
        const bool hetero = n_delays > 1;
        if (hetero) {
            if (n_delays != n_syn)
                throw std::runtime_error("b200: number of delays does not match number of synapses");
            dsteps.resize(n_syn);
            for (size_t i = 0; i < n_syn; ++i) dsteps[i] = (int)(real_delays[i] / dt + 0.5);
        } else if (n_delays == 1) {
            scalar_delay = (int)(real_delays[0] / dt + 0.5);
        }
        // distinct delays
        bin_delay.clear();
        if (hetero) {
            int mx = 0;
            for (size_t i = 0; i < n_syn; ++i) {
                if (dsteps[i] < 0) throw std::runtime_error("b200: negative synaptic delay");
                mx = std::max(mx, dsteps[i]);
            }
            std::vector<char> seen(mx + 1, 0);
            for (size_t i = 0; i < n_syn; ++i) seen[dsteps[i]] = 1;
            std::vector<int> bin_of(mx + 1, -1);
            for (int d = 0; d <= mx; ++d)
                if (seen[d]) { bin_of[d] = (int)bin_delay.size(); bin_delay.push_back(d); }
            for (size_t i = 0; i < n_syn; ++i) dsteps[i] = bin_of[dsteps[i]];   // now bin index
        } else {
            bin_delay.push_back(scalar_delay);
        }
        nbins = (int)bin_delay.size();
        max_delay = bin_delay.empty() ? 0 : bin_delay.back();
        // FORWARD layout (counted pathways whose delays are all >= 1 step, at most 32 distinct):
        // CSR by (source, delay bin) with the bin packed into the top 5 bits of every target
        // word.  A spike is then delivered ONCE, one step after it was emitted -- its whole row
        // is one contiguous piece of the index stream -- into the hit counters of the step in
        // which each synapse is due (`hits[(t_spike + delay) % (max_delay + 1)][target]`), instead
        // of once per delay bin from a compacted list of the right age: 20x fewer, 20x longer
        // rows for Brunel's 20 delays, and no compaction of spike lists at all.
        forward = false;
        hits_slots = 2;
        if (counted_n > 0 && allow_forward() && nbins >= 1 && nbins <= 32 && bin_delay.front() >= 1 && targets) {
            forward = true;
            for (size_t i = 0; i < n_syn && forward; ++i)
                if (targets[i] < 0 || targets[i] >= (1 << 27)) forward = false;
        }
        if (forward) {
            std::vector<int> rowptr, csr_target;
            n_owned = build_forward_csr(nsrc, nbins, spikes_start, srcs, targets, hetero ? dsteps.data() : nullptr,
                                        n_syn, owned, rowptr, csr_target);
            identity = false;
            std::vector<int> bin_info(bin_delay.begin(), bin_delay.end());
            bin_info.resize(2 * nbins, 0);
            d_bin_delay = (int*)dev_alloc(std::max<size_t>(1, 2 * nbins) * sizeof(int));
            d_rowptr = (int*)dev_alloc(rowptr.size() * sizeof(int));
            d_csr_target = (int*)dev_alloc(std::max<size_t>(1, n_owned) * sizeof(int));
            B200_CUDA(cudaMemcpy(d_bin_delay, bin_info.data(), 2 * nbins * sizeof(int), cudaMemcpyHostToDevice));
            B200_CUDA(cudaMemcpy(d_rowptr, rowptr.data(), rowptr.size() * sizeof(int), cudaMemcpyHostToDevice));
            if (n_owned)
                B200_CUDA(cudaMemcpy(d_csr_target, csr_target.data(), n_owned * sizeof(int), cudaMemcpyHostToDevice));
            if (!d_events) {
                d_events = (unsigned long long*)dev_alloc(sizeof(unsigned long long));
                B200_CUDA(cudaMemset(d_events, 0, sizeof(unsigned long long)));
            }
            if (!d_tickets) d_tickets = (unsigned int*)dev_alloc(2 * sizeof(unsigned int));
            hits_slots = max_delay + 1;
            // only the previous step's list is ever read: straight from the thresholder's segments
            if (es) es->require(1, 1);
            prepared = true;
            state().prepare_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - _t0).count();
            return;
        }
        // counting sort by (bin, source)
        const size_t nrows = (size_t)nbins * (nsrc + 1);
        std::vector<int> rowptr(nrows + 1, 0);
        // row r = bin*(nsrc+1) + src ; the extra row per bin keeps rowptr[bin][nsrc] addressable
        n_owned = 0;
        for (size_t i = 0; i < n_syn; ++i) {
            const int s = srcs[i] - spikes_start;
            if (s < 0 || s >= nsrc) throw std::runtime_error("b200: synapse source outside pathway source range");
            if (!owned(i)) continue;
            const size_t r = (size_t)(hetero ? dsteps[i] : 0) * (nsrc + 1) + s;
            rowptr[r + 1]++;
            n_owned++;
        }
        for (size_t r = 0; r < nrows; ++r) rowptr[r + 1] += rowptr[r];
        std::vector<int> syn_ids(n_owned), csr_target(n_owned);
        {
            std::vector<int> cursor(rowptr.begin(), rowptr.end() - 1);
            for (size_t i = 0; i < n_syn; ++i) {
                if (!owned(i)) continue;
                const int s = srcs[i] - spikes_start;
                const size_t r = (size_t)(hetero ? dsteps[i] : 0) * (nsrc + 1) + s;
                syn_ids[cursor[r]++] = (int)i;
            }
        }
        identity = n_owned == n_syn;
        for (size_t k = 0; k < n_owned; ++k) {
            if (syn_ids[k] != (int)k) identity = false;
            csr_target[k] = targets ? targets[syn_ids[k]] : 0;
        }
        // longest row of every bin: the propagation kernel cuts rows into chunks of equal length
        std::vector<int> bin_info(bin_delay.begin(), bin_delay.end());
        for (int b = 0; b < nbins; ++b) {
            int mx = 0;
            const int* rp = rowptr.data() + (size_t)b * (nsrc + 1);
            for (int s = 0; s < nsrc; ++s) mx = std::max(mx, rp[s + 1] - rp[s]);
            bin_info.push_back(mx);
        }
        d_bin_delay = (int*)dev_alloc(std::max<size_t>(1, 2 * nbins) * sizeof(int));
        d_rowptr = (int*)dev_alloc((nrows + 1) * sizeof(int));
        d_syn_ids = (int*)dev_alloc(std::max<size_t>(1, n_owned) * sizeof(int));
        d_csr_target = (int*)dev_alloc(std::max<size_t>(1, n_owned) * sizeof(int));
        if (nbins) B200_CUDA(cudaMemcpy(d_bin_delay, bin_info.data(), 2 * nbins * sizeof(int), cudaMemcpyHostToDevice));
        B200_CUDA(cudaMemcpy(d_rowptr, rowptr.data(), (nrows + 1) * sizeof(int), cudaMemcpyHostToDevice));
        if (n_owned) {
            B200_CUDA(cudaMemcpy(d_syn_ids, syn_ids.data(), n_owned * sizeof(int), cudaMemcpyHostToDevice));
            B200_CUDA(cudaMemcpy(d_csr_target, csr_target.data(), n_owned * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (!d_events) {
            d_events = (unsigned long long*)dev_alloc(sizeof(unsigned long long));
            B200_CUDA(cudaMemset(d_events, 0, sizeof(unsigned long long)));
        }
        if (!d_tickets) d_tickets = (unsigned int*)dev_alloc(2 * sizeof(unsigned int));
        B200_CUDA(cudaMemset(d_tickets, 0, 2 * sizeof(unsigned int)));
        if (es) es->require(bin_delay.empty() ? 0 : bin_delay.front(), max_delay);
        prepared = true;
        state().prepare_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - _t0).count();
    }

    // counted pathways: two zeroed counter arrays (by step parity) over the target group
    void ensure_hits(int n) {
        if (n <= 0 || (d_hits && hits_n == n && hits_alloc_slots == hits_slots)) return;
        dev_free(d_hits);
        hits_n = n;
        hits_alloc_slots = hits_slots;
        d_hits = (int*)dev_alloc((size_t)hits_slots * (size_t)n * sizeof(int));
        B200_CUDA(cudaMemset(d_hits, 0, (size_t)hits_slots * (size_t)n * sizeof(int)));
    }

    PathwayDev view() const {
        PathwayDev v;
        v.nsrc = spikes_stop - spikes_start;
        v.src_start = spikes_start;
        v.nbins = nbins;
        v.identity = identity ? 1 : 0;
        const int seg = state().all_delayed ? 1 : 0;
        v.seg_delay = (!bin_delay.empty() && bin_delay.front() == seg) ? seg : -1;
        if (forward) v.seg_delay = 1;
        v.hits = d_hits;
        v.hits_n = hits_n;
        v.hits_slots = hits_slots;
        v.forward = forward ? 1 : 0;
        v.tileptr = d_tileptr;
        v.tile_stride = tile_stride;
        v.bin_delay = d_bin_delay;
        v.bin_maxlen = d_bin_delay + nbins;
        v.rowptr = d_rowptr;
        v.syn_ids = d_syn_ids;
        v.csr_target = d_csr_target;
        v.es = es ? es->id : -1;
        v.events = d_events;
        v.tickets = d_tickets;
        return v;
    }

    unsigned long long events_delivered() const {
        unsigned long long e = 0;
        if (d_events) cudaMemcpy(&e, d_events, sizeof(e), cudaMemcpyDeviceToHost);
        return e;
    }
};

// ---------------------------------------------------------------------------------------------
// TargetIndex: CSR by target element for summed variables (summed_variable.cpp:20-27 adds the
// synapses' values in synapse-index order; a counting sort by target keeps that order per row).
// ---------------------------------------------------------------------------------------------
class TargetIndex {
public:
    int n_targets = 0;
    int* d_rowptr = nullptr;
    int* d_syn_ids = nullptr;
    bool prepared = false;

    ~TargetIndex() { release(); }
    void release() {
        dev_free(d_rowptr); dev_free(d_syn_ids);
        d_rowptr = d_syn_ids = nullptr;
    }
    // index[i]: absolute target element of synapse i; rows cover [target_start, target_start + size)
    void prepare(const int32_t* index, size_t n_syn, int target_start, int target_size) {
        runtime_init();
        release();
        if (n_syn >= (size_t)INT32_MAX) throw std::runtime_error("b200: more than 2^31-1 synapses");
        n_targets = target_size;
        std::vector<int> rowptr((size_t)target_size + 1, 0);
        size_t kept = 0;
        for (size_t i = 0; i < n_syn; ++i) {
            const int t = index[i] - target_start;
            if (t < 0 || t >= target_size) continue;
            rowptr[(size_t)t + 1]++;
            kept++;
        }
        for (int t = 0; t < target_size; ++t) rowptr[(size_t)t + 1] += rowptr[t];
        std::vector<int> syn_ids(kept);
        {
            std::vector<int> cursor(rowptr.begin(), rowptr.end() - 1);
            for (size_t i = 0; i < n_syn; ++i) {
                const int t = index[i] - target_start;
                if (t < 0 || t >= target_size) continue;
                syn_ids[cursor[t]++] = (int)i;
            }
        }
        d_rowptr = (int*)dev_alloc(rowptr.size() * sizeof(int));
        d_syn_ids = (int*)dev_alloc(std::max<size_t>(1, kept) * sizeof(int));
        B200_CUDA(cudaMemcpy(d_rowptr, rowptr.data(), rowptr.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (kept) B200_CUDA(cudaMemcpy(d_syn_ids, syn_ids.data(), kept * sizeof(int), cudaMemcpyHostToDevice));
        prepared = true;
    }
    TargetIndexDev view() const {
        TargetIndexDev v;
        v.n_targets = n_targets;
        v.rowptr = d_rowptr;
        v.syn_ids = d_syn_ids;
        return v;
    }
};

}  // namespace b200
