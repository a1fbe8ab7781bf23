// CPU check of b200::Pathway::build_forward_csr (csrc/b200_host.h): the (source, delay bin)
// layout of counted pathways with delays >= 1 step.  Random synapses (unsorted, with a partition
// filter like the multi-GPU one); every kept synapse must appear exactly once, in the row of its
// source, inside the bin range of its delay, with the bin packed into the top 5 bits, and the
// synapses of one (source, bin) must keep their original relative order.  Prints "OK <n kept>".
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include "b200_host.h"

int main() {
    unsigned long long z = 12345;
    auto rnd = [&]() { z = z * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned int)(z >> 33); };
    for (int trial = 0; trial < 20; ++trial) {
        const int nsrc = 1 + rnd() % 50, ntgt = 1 + rnd() % 200, nbins = 1 + rnd() % 32, start = rnd() % 7;
        const size_t n = rnd() % 5000;
        std::vector<int> srcs(n), tgts(n), bins(n);
        for (size_t i = 0; i < n; ++i) { srcs[i] = start + rnd() % nsrc; tgts[i] = rnd() % ntgt; bins[i] = rnd() % nbins; }
        const int lo = ntgt / 4, hi = ntgt - ntgt / 4;                 // "owned" targets
        auto owned = [&](size_t i) { return trial % 2 == 0 || (tgts[i] >= lo && tgts[i] < hi); };
        std::vector<int> rowptr, csr;
        const size_t kept = b200::Pathway::build_forward_csr(nsrc, nbins, start, srcs.data(), tgts.data(),
                                                             nbins > 1 ? bins.data() : nullptr, n, owned, rowptr, csr);
        size_t expect = 0;
        std::map<std::pair<int, int>, std::vector<int>> rows;      // (source, bin) -> targets in order
        for (size_t i = 0; i < n; ++i)
            if (owned(i)) { expect++; rows[{srcs[i] - start, nbins > 1 ? bins[i] : 0}].push_back(tgts[i]); }
        if (kept != expect || csr.size() != kept || rowptr.size() != (size_t)nsrc * (nbins + 1) + 1) { printf("FAIL sizes\n"); return 1; }
        int prev_end = 0;
        for (int s = 0; s < nsrc; ++s) {
            const int* rp = rowptr.data() + (size_t)s * (nbins + 1);
            if (rp[0] != prev_end) { printf("FAIL rows not contiguous\n"); return 1; }
            for (int b = 0; b < nbins; ++b) {
                const std::vector<int>& want = rows[{s, b}];
                if (rp[b + 1] - rp[b] != (int)want.size()) { printf("FAIL row length\n"); return 1; }
                for (size_t k = 0; k < want.size(); ++k) {
                    const unsigned int w = (unsigned int)csr[rp[b] + k];
                    if ((int)(w >> 27) != b || (int)(w & 0x7ffffffu) != want[k]) { printf("FAIL entry\n"); return 1; }
                }
            }
            prev_end = rp[nbins];
        }
        if (prev_end != (int)kept) { printf("FAIL total\n"); return 1; }
    }
    printf("OK\n");
    return 0;
}
