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
// (bid, nb) is the VIRTUAL grid a code object distributes its work over: the whole grid for the
// per-element code objects (they share one partition), a sub-range of CTAs when several
// independent code objects of one phase run side by side.  (gbid, gnb) is the real grid.
struct Ctx {
    int bid;
    int nb;
    int gbid;
    int gnb;
    int rank;
    int world;
};

constexpr int kMaxRanks = 8;           // GPUs of one box that can share a network
constexpr int kMaxSegments = kMaxRanks * kBlock;   // spike-list segments of one event space

// Device view of one event space ("spike ring").  A step's spike list exists in two forms:
//   * segments: CTA b of rank r writes the ids of ITS neurons that fired (ascending) at
//     ids[slot][seg_start[r*nb+b] ...] and their number at cnt[slot][r*nb + b]; the
//     concatenation in segment order is the ascending `_spikespace` of the reference
//     (threshold.cpp:24-31).  Produced without any inter-CTA communication.  Every word carries
//     the tag of its time step (ids: tag << 32 | id, counts: tag16 << 16 | count), so on several
//     GPUs the same stores simply also go to every peer's ring over NVLink and a reader spins
//     on the word it needs until the tag matches -- flag and data travel in one word, no fence
//     or ordering between stores is required on either side.
//   * compact: ids [0, count) + count at [N], the reference layout, built `lag` steps later by
//     `compact_segments` for the consumers that look back in time (synaptic delays).
struct EventSpaceDev {
    unsigned long long* ids;  // [slots][N]
    int32_t* cnt;             // [slots][nseg]
    int32_t* compact;         // [slots][N + 1]
    const int32_t* seg_start; // [nseg + 1] first neuron of every segment (absolute id)
    int slots;
    int N;
    int nseg;                 // world * nb
    int lag;                  // compaction of step s happens during step s + lag
    int id;                   // index of this event space (tag of the cached view)
    int need_compact;         // 0: nobody reads the compact form (no delayed / serial pathway)
    int rank_lo, rank_hi;     // neurons owned by this rank
    // multi-GPU: the peers' rings (CUDA IPC mapped)
    unsigned long long* peer_ids[kMaxRanks];
    int32_t* peer_cnt[kMaxRanks];
};

// Device view of a synaptic pathway (built by b200_host.h: Pathway::prepare)
struct PathwayDev {
    int nsrc;                 // number of source neurons (source.stop - source.start)
    int src_start;            // first source id in the parent group (spikequeue.h:97,162)
    int nbins;                // distinct integer delays
    int identity;             // 1: csr slot k == synapse index k (no indirection needed)
    int seg_delay;            // delay (0 or 1 step) whose spike list is read straight from the
                              // thresholder's segments, -1: none (older lists: compacted form)
    int* hits;                // counted pathways: [2][hits_n] events per target, by step parity
    int hits_n;
    int hits_slots;           // 2 (by step parity), or max_delay + 1 in the forward layout
    int forward;              // 1: CSR by (source, delay bin), bin in the top 5 bits of csr_target,
                              // rowptr[nsrc][nbins + 1]; a spike is delivered once, one step later
    const int* tileptr;       // dense rows (b200_tiles.cuh): [nbins][nsrc + 1][grid + 1] first slot of
                              // the row at or after the first target of every CTA's block; 0: none
    int tile_stride;          // ints between the per-warp counter arrays in shared memory
    const int* bin_delay;     // [nbins] delay in steps, ascending
    const int* bin_maxlen;    // [nbins] length of the longest row of the bin
    const int* rowptr;        // [nbins*(nsrc+1)+1] slot offsets
    const int* syn_ids;       // [S] synapse index per slot (sorted by delay, source, index)
    const int* csr_target;    // [S] the non-source end of the synapse, packed in slot order
    int es;                   // (unused on the device; kept for debugging)
    unsigned long long* events;   // number of delivered synaptic events (for the metric)
    unsigned int* tickets;        // [2] work counters of heavy steps (by step parity)
};

// Device view of a by-target index of a Synapses object (summed variables): row t lists the
// synapses (ascending index) whose target element is t + target_start
struct TargetIndexDev {
    int n_targets;
    const int* rowptr;        // [n_targets + 1]
    const int* syn_ids;       // [number of synapses with a target in range]
};

// Control block shared by host and the persistent kernel
struct Control {
    unsigned long long barrier;     // grid barrier arrival counter
    int stop;                       // != 0: leave the step loop after the current step
    int steps_done;                 // steps completed by the last launch
    int overflow;                   // a monitor buffer is (nearly) full: host must grow it
    int error;                      // != 0: device-side failure (1: peer wait timed out)
    unsigned long long poll_cycles; // multi-GPU diagnostics (one sampled thread): cycles spent
    unsigned long long fence_cycles;//   spinning on peers' ready flags / in the acquire fence
    unsigned long long polls;       //   number of waits sampled
};

}  // namespace b200
