// b200_types.h -- plain structs shared by host code (g++) and device code (nvcc) of the b200 device.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace b200 {

constexpr int kBlock = 512;            // threads per CTA for every generated kernel
constexpr int kWarps = kBlock / 32;
constexpr int kMaxOwnedIters = 64;     // per-lane elements of an owned slice (bit mask width)

// Value of one Brian clock for the current step (brianlib/clocks.h:34-38: t = timestep*dt)
struct Clk {
    double t;
    double dt;
    int64_t timestep;
};

// Virtual CTA coordinates (identical in per-code-object kernels and in the persistent kernel)
struct Ctx {
    int bid;
    int nb;
};

// Device view of a synaptic pathway (built by b200_host.h: Pathway::prepare)
struct PathwayDev {
    int nsrc;                 // number of source neurons (source.stop - source.start)
    int src_start;            // first source id in the parent group (spikequeue.h:97,162)
    int nbins;                // distinct integer delays
    int identity;             // 1: csr slot k == synapse index k (no indirection needed)
    const int* bin_delay;     // [nbins] delay in steps, ascending
    const int* rowptr;        // [nbins*(nsrc+1)+1] slot offsets
    const int* syn_ids;       // [S] synapse index per slot (sorted by delay, source, index)
    const int* csr_target;    // [S] the non-source end of the synapse, packed in slot order
    const int32_t* ring;      // spike ring of the source event space
    int ring_slots;
    int ring_stride;          // N_group + 1
    unsigned long long* events;   // number of delivered synaptic events (for the metric)
};

// Control block shared by host and the persistent kernel
struct Control {
    unsigned long long barrier;     // grid barrier arrival counter
    int stop;                       // != 0: leave the step loop after the current step
    int steps_done;                 // steps completed by the last launch
    int overflow;                   // a monitor buffer is (nearly) full: host must grow it
    int pad;
};

}  // namespace b200
