// b200_tiles.cuh -- target tiles of a counted pathway with dense rows (templates/synapses.cu,
// apply pass): every (delay bin, source) row of the CSR is cut at the boundaries of the CTAs'
// blocks of the target group, so that the CTA that owns a block of targets can read, for every
// spiking source, exactly the piece of the packed index stream that points into its block.
//
// Reference semantics served: brian2/devices/cpp_standalone/templates/synapses.cpp:20-49 (the
// effect of every queued synapse is applied to its target) -- here by the owner of the target.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "b200_host.h"

namespace b200 {

// rows must list their targets in strictly ascending order (what connect() produces; a row with
// repeated or unsorted targets -- multapses, several connect calls -- keeps the scatter path)
__global__ void tiles_check_rows(const int* rowptr, const int* csr_target, long long nrows, int* unsorted) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (long long r = warp; r < nrows; r += nwarps) {
        const int beg = rowptr[r], end = rowptr[r + 1];
        for (int k = beg + lane; k + 1 < end; k += 32)
            if (csr_target[k] >= csr_target[k + 1]) *unsorted = 1;
    }
}

// tileptr[r * (ntiles + 1) + b] = first slot of row r whose target is >= tile_start[b]
__global__ void tiles_cut_rows(const int* rowptr, const int* csr_target, long long nrows, int ntiles,
                               const int* tile_start, int* tileptr) {
    const long long total = nrows * (long long)(ntiles + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (ntiles + 1);
        const int b = (int)(i - r * (ntiles + 1));
        const int x = tile_start[b];
        int lo = rowptr[r], hi = rowptr[r + 1];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (csr_target[mid] < x) lo = mid + 1; else hi = mid;
        }
        tileptr[i] = lo;
    }
}

constexpr int kMaxTile = 1024;                       // targets per CTA block that fit the counters
constexpr size_t kTileSmem = (size_t)kWarps * kMaxTile * sizeof(int);

// Could this pathway be served by tiles on a grid of about `grid_guess` CTAs?  (Decides the
// dynamic shared memory of the step kernels before the grid size is final.)
inline bool tiles_candidate(const Pathway& pw, int grid_guess) {
    if (!pw.prepared || pw.forward || pw.tile_n <= 0 || pw.nbins <= 0 || pw.n_owned == 0) return false;
    const double nrows = (double)pw.nbins * (double)std::max(1, pw.spikes_stop - pw.spikes_start);
    return (double)pw.n_owned / nrows >= 8.0 * (double)grid_guess;
}

// (Re)build the tile table of `pw` for `grid` CTAs; leaves pw.d_tileptr == nullptr if the
// pathway does not qualify.  `n_target` = size of the target group (== hits_n).
inline void tiles_build(Pathway& pw, int grid, size_t dyn_smem) {
    dev_free(pw.d_tileptr);
    pw.d_tileptr = nullptr;
    pw.tile_stride = 0;
    if (!tiles_candidate(pw, grid) || dyn_smem < kTileSmem) return;
    RuntimeState& st = state();
    const int nsrc = pw.spikes_stop - pw.spikes_start;
    const long long nrows = (long long)pw.nbins * (nsrc + 1) - 0;
    // tile boundaries = the CTAs' blocks of the target group (same arithmetic as owned_cta)
    int64_t lo, hi;
    EventSpace::rank_range_host(pw.tile_n, st.rank, st.world, lo, hi);
    std::vector<int> start(grid + 1);
    int tmax = 0;
    for (int b = 0; b <= grid; ++b)
        start[b] = (int)EventSpace::warp_first_host(lo, hi, (int64_t)b * kWarps, (int64_t)grid * kWarps);
    for (int b = 0; b < grid; ++b) tmax = std::max(tmax, start[b + 1] - start[b]);
    start[0] = INT32_MIN;           // everything below the rank's block belongs to "before tile 0"
    const size_t table = (size_t)nrows * (size_t)(grid + 1) * sizeof(int);
    if (tmax > kMaxTile || table > ((size_t)256 << 20)) return;
    int* d_flag = (int*)dev_alloc(sizeof(int));
    B200_CUDA(cudaMemset(d_flag, 0, sizeof(int)));
    tiles_check_rows<<<296, 512, 0, st.stream>>>(pw.d_rowptr, pw.d_csr_target, nrows, d_flag);
    int unsorted = 0;
    B200_CUDA(cudaMemcpyAsync(&unsorted, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st.stream));
    B200_CUDA(cudaStreamSynchronize(st.stream));
    dev_free(d_flag);
    if (unsorted) return;
    int* d_start = (int*)dev_alloc((grid + 1) * sizeof(int));
    B200_CUDA(cudaMemcpy(d_start, start.data(), (grid + 1) * sizeof(int), cudaMemcpyHostToDevice));
    pw.d_tileptr = (int*)dev_alloc(table);
    tiles_cut_rows<<<296, 512, 0, st.stream>>>(pw.d_rowptr, pw.d_csr_target, nrows, grid, d_start, pw.d_tileptr);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(st.stream));
    dev_free(d_start);
    pw.tile_stride = ((tmax + 31) & ~31) + 1;
    if ((size_t)kWarps * pw.tile_stride * sizeof(int) > dyn_smem) pw.tile_stride = tmax;
    pw.tiles_grid = grid;
}

}  // namespace b200
