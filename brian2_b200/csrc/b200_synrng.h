// b200_synrng.h -- host-side random numbers for per-synapse initialisation under "sharded
// construction" (prefs.devices.b200.construction = 'sharded').
//
// The reference fills `S.delay = '... rand() ...'` from its one sequential mt19937 stream
// (brian2/devices/cpp_standalone/templates/objects.cpp:426-479, group_variable_set*.cpp), so the
// value of a synapse depends on how many synapses precede it -- which differs between ranks once
// every rank only holds the synapses of its own postsynaptic neurons.  Here the draw of a synapse
// is a pure function of (seed, code object, presynaptic index, postsynaptic index, occurrence,
// call number): Philox4x32-10 with that tuple as key/counter.  The same network therefore gets
// the same delays and weights on 1, 2, 4 or 8 GPUs.
#pragma once
#include <stdint.h>
#include <cmath>
#include <vector>

#include "b200_host.h"

namespace b200 {

inline void philox4x32_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One generator per code object; the loop over the synapses is sequential (group_variable_set*.cpp
// iterates _idx = 0 .. N-1), which is all `occurrence` (synapses with equal (pre, post) that
// follow each other, e.g. connect(..., n=2)) relies on.
struct SynapseRng {
    uint32_t stream;
    long last_idx = -1;
    int32_t last_pre = -1, last_post = -1;
    uint32_t occurrence = 0, call = 0;
    double spare = 0.0;
    bool has_spare = false;

    explicit SynapseRng(uint32_t stream_) : stream(stream_) {}

    void select(long idx, const std::vector<int32_t>& pre, const std::vector<int32_t>& post) {
        if (idx == last_idx) return;
        const int32_t p = pre[(size_t)idx], q = post[(size_t)idx];
        occurrence = (idx == last_idx + 1 && p == last_pre && q == last_post) ? occurrence + 1 : 0;
        last_idx = idx; last_pre = p; last_post = q;
        call = 0;
        has_spare = false;
    }
    double uniform(long idx, const std::vector<int32_t>& pre, const std::vector<int32_t>& post) {
        ensure_seed();
        select(idx, pre, post);
        const unsigned long long seed = state().seed;
        uint32_t o[4];
        philox4x32_host((uint32_t)last_pre, (uint32_t)last_post, occurrence, call++,
                        (uint32_t)seed ^ (stream * 0x9E3779B9u), (uint32_t)(seed >> 32) ^ stream, o);
        // 53 random bits, the reference's recipe (objects.cpp:448-453)
        const double a = (double)(o[0] >> 5), b = (double)(o[1] >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    // polar Box-Muller with a cached second value per synapse (objects.cpp:455-478)
    double normal(long idx, const std::vector<int32_t>& pre, const std::vector<int32_t>& post) {
        select(idx, pre, post);
        if (has_spare) { has_spare = false; return spare; }
        double x1, x2, r2;
        do {
            x1 = 2.0 * uniform(idx, pre, post) - 1.0;
            x2 = 2.0 * uniform(idx, pre, post) - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = std::sqrt(-2.0 * std::log(r2) / r2);
        spare = f * x1;
        has_spare = true;
        return f * x2;
    }
};

}  // namespace b200
