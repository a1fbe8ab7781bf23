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
#include <cstdio>
#include <cstring>
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
    unsigned long long launches = 0;   // kernels launched inside run loops
    double upload_seconds = 0.0, download_seconds = 0.0;
    size_t h2d_bytes = 0, d2h_bytes = 0;
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
    const char* lr = getenv("LOCAL_RANK");
    const char* rk = getenv("RANK");
    const char* ws = getenv("WORLD_SIZE");
    const char* ng = getenv("B200_MULTI_GPU");
    if (ng && atoi(ng) > 0 && ws && atoi(ws) > 1) {
        s.world = atoi(ws);
        s.rank = rk ? atoi(rk) : 0;
    }
    s.device = lr ? atoi(lr) % ndev : 0;
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
// Event ring: the last `slots` spike lists of one event space, [slots][stride] int32, list s is
// stored at slot (timestep % slots), ids ascending, count in the last element (same layout as
// `_spikespace`, threshold.cpp:24-31).
// ---------------------------------------------------------------------------------------------
struct EventRing {
    int32_t* dev = nullptr;
    int slots = 0;
    int stride = 0;            // N + 1
    int required = 1;          // max delay + 1 over all pathways reading this event space
    unsigned long long* scan_ws = nullptr;   // look-back workspace (kBlock entries)

    void require(int nslots) { required = std::max(required, nslots); }

    // (re)allocate; keeps the history of the last min(slots, old slots) steps before `timestep`
    void ensure(int stride_, int64_t timestep) {
        if (!scan_ws) {
            scan_ws = (unsigned long long*)dev_alloc(kBlock * sizeof(unsigned long long));
            B200_CUDA(cudaMemset(scan_ws, 0, kBlock * sizeof(unsigned long long)));
        }
        if (dev && slots >= required && stride == stride_) return;
        const int new_slots = required;
        int32_t* nd = (int32_t*)dev_alloc((size_t)new_slots * stride_ * sizeof(int32_t));
        B200_CUDA(cudaMemset(nd, 0, (size_t)new_slots * stride_ * sizeof(int32_t)));
        if (dev && stride == stride_) {
            // the old ring holds steps timestep-1 ... timestep-slots
            for (int back = 1; back <= slots && back < new_slots; ++back) {
                const int64_t s = timestep - back;
                int64_t os = s % slots, ns = s % new_slots;
                if (os < 0) os += slots;
                if (ns < 0) ns += new_slots;
                B200_CUDA(cudaMemcpy(nd + ns * (size_t)stride_, dev + os * (size_t)stride,
                                     stride * sizeof(int32_t), cudaMemcpyDeviceToDevice));
            }
        }
        dev_free(dev);
        dev = nd;
        slots = new_slots;
        stride = stride_;
    }
    int32_t* slot_ptr(int64_t timestep) const {
        int64_t s = timestep % slots;
        if (s < 0) s += slots;
        return dev + s * (size_t)stride;
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
    std::vector<int> bin_delay;
    // device storage
    int* d_bin_delay = nullptr;
    int* d_rowptr = nullptr;
    int* d_syn_ids = nullptr;
    int* d_csr_target = nullptr;
    unsigned long long* d_events = nullptr;
    EventRing* ring = nullptr;

    Pathway(std::vector<int>& _sources, int _spikes_start, int _spikes_stop)
        : sources(_sources), spikes_start(_spikes_start), spikes_stop(_spikes_stop) {}

    ~Pathway() { release(); }

    void release() {
        dev_free(d_bin_delay); dev_free(d_rowptr); dev_free(d_syn_ids); dev_free(d_csr_target);
        d_bin_delay = d_rowptr = d_syn_ids = d_csr_target = nullptr;
    }

    // Build the delay-binned CSR.  Delay rounding as CSpikeQueue::prepare (spikequeue.h:90):
    // steps = (int)(delay/dt + 0.5); n_delays == 1 means one delay for all synapses (:104).
    // Slot order = (delay bin asc, source asc, synapse index asc); walking the bins from the
    // largest delay to the smallest reproduces the reference's delivery order within a step
    // (entries pushed earlier sit first in a bucket, spikequeue.h:157-190).
    template <typename scalar>
    void prepare(int n_source, int n_target, const scalar* real_delays, size_t n_delays,
                 const int* srcs, const int* targets, size_t n_syn, double dt, EventRing* ring_) {
        runtime_init();
        release();
        Nsource = n_source;
        Ntarget = n_target;
        n_synapses = n_syn;
        ring = ring_;
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
        // counting sort by (bin, source)
        const size_t nrows = (size_t)nbins * (nsrc + 1);
        std::vector<int> rowptr(nrows + 1, 0);
        // row r = bin*(nsrc+1) + src ; the extra row per bin keeps rowptr[bin][nsrc] addressable
        for (size_t i = 0; i < n_syn; ++i) {
            const int s = srcs[i] - spikes_start;
            if (s < 0 || s >= nsrc) throw std::runtime_error("b200: synapse source outside pathway source range");
            const size_t r = (size_t)(hetero ? dsteps[i] : 0) * (nsrc + 1) + s;
            rowptr[r + 1]++;
        }
        for (size_t r = 0; r < nrows; ++r) rowptr[r + 1] += rowptr[r];
        std::vector<int> syn_ids(n_syn), csr_target(n_syn);
        {
            std::vector<int> cursor(rowptr.begin(), rowptr.end() - 1);
            for (size_t i = 0; i < n_syn; ++i) {
                const int s = srcs[i] - spikes_start;
                const size_t r = (size_t)(hetero ? dsteps[i] : 0) * (nsrc + 1) + s;
                syn_ids[cursor[r]++] = (int)i;
            }
        }
        identity = true;
        for (size_t k = 0; k < n_syn; ++k) {
            if (syn_ids[k] != (int)k) identity = false;
            csr_target[k] = targets ? targets[syn_ids[k]] : 0;
        }
        d_bin_delay = (int*)dev_alloc(std::max<size_t>(1, nbins) * sizeof(int));
        d_rowptr = (int*)dev_alloc((nrows + 1) * sizeof(int));
        d_syn_ids = (int*)dev_alloc(std::max<size_t>(1, n_syn) * sizeof(int));
        d_csr_target = (int*)dev_alloc(std::max<size_t>(1, n_syn) * sizeof(int));
        if (nbins) B200_CUDA(cudaMemcpy(d_bin_delay, bin_delay.data(), nbins * sizeof(int), cudaMemcpyHostToDevice));
        B200_CUDA(cudaMemcpy(d_rowptr, rowptr.data(), (nrows + 1) * sizeof(int), cudaMemcpyHostToDevice));
        if (n_syn) {
            B200_CUDA(cudaMemcpy(d_syn_ids, syn_ids.data(), n_syn * sizeof(int), cudaMemcpyHostToDevice));
            B200_CUDA(cudaMemcpy(d_csr_target, csr_target.data(), n_syn * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (!d_events) {
            d_events = (unsigned long long*)dev_alloc(sizeof(unsigned long long));
            B200_CUDA(cudaMemset(d_events, 0, sizeof(unsigned long long)));
        }
        if (ring) ring->require(max_delay + 1);
        prepared = true;
    }

    PathwayDev view() const {
        PathwayDev v;
        v.nsrc = spikes_stop - spikes_start;
        v.src_start = spikes_start;
        v.nbins = nbins;
        v.identity = identity ? 1 : 0;
        v.bin_delay = d_bin_delay;
        v.rowptr = d_rowptr;
        v.syn_ids = d_syn_ids;
        v.csr_target = d_csr_target;
        v.ring = ring ? ring->dev : nullptr;
        v.ring_slots = ring ? ring->slots : 1;
        v.ring_stride = ring ? ring->stride : 1;
        v.events = d_events;
        return v;
    }

    unsigned long long events_delivered() const {
        unsigned long long e = 0;
        if (d_events) cudaMemcpy(&e, d_events, sizeof(e), cudaMemcpyDeviceToHost);
        return e;
    }
};

}  // namespace b200
