// CPU check of b200::SynapseRng (csrc/b200_synrng.h): the draw of a synapse is a pure function of
// (seed, stream, pre, post, occurrence, call number) -- independent of where the synapse sits in
// a rank's local array -- and the draws are uniform / standard normal.  Prints "OK".
#include <cmath>
#include <cstdio>
#include <map>
#include <vector>
#include "b200_synrng.h"

int main() {
    b200::state().seed = 1234;
    b200::state().seeded = true;
    // a "global" synapse list in (pre, post) order, with one duplicated pair
    std::vector<int32_t> pre, post;
    for (int i = 0; i < 300; ++i)
        for (int j = (i * 7) % 5; j < 400; j += 5) { pre.push_back(i); post.push_back(j); }
    pre.insert(pre.begin() + 1000, pre[999]);
    post.insert(post.begin() + 1000, post[999]);
    b200::SynapseRng whole(77u);
    std::vector<double> u(pre.size()), g(pre.size());
    for (size_t k = 0; k < pre.size(); ++k) u[k] = whole.uniform((long)k, pre, post);
    b200::SynapseRng whole_n(78u);
    for (size_t k = 0; k < pre.size(); ++k) g[k] = whole_n.normal((long)k, pre, post);
    if (u[999] == u[1000]) { printf("FAIL duplicate pair drew the same number\n"); return 1; }
    // two "ranks": synapses split by postsynaptic neuron, each with its own local arrays
    for (int rank = 0; rank < 2; ++rank) {
        std::vector<int32_t> lp, lq;
        std::vector<size_t> global;
        for (size_t k = 0; k < pre.size(); ++k)
            if ((post[k] < 200) == (rank == 0)) { lp.push_back(pre[k]); lq.push_back(post[k]); global.push_back(k); }
        b200::SynapseRng part(77u), part_n(78u);
        for (size_t k = 0; k < lp.size(); ++k) {
            if (part.uniform((long)k, lp, lq) != u[global[k]]) { printf("FAIL uniform differs on rank %d\n", rank); return 1; }
            if (part_n.normal((long)k, lp, lq) != g[global[k]]) { printf("FAIL normal differs on rank %d\n", rank); return 1; }
        }
    }
    double mu = 0, var = 0, gm = 0, gv = 0;
    for (size_t k = 0; k < u.size(); ++k) { mu += u[k]; gm += g[k]; }
    mu /= u.size(); gm /= u.size();
    for (size_t k = 0; k < u.size(); ++k) { var += (u[k] - mu) * (u[k] - mu); gv += (g[k] - gm) * (g[k] - gm); }
    var /= u.size(); gv /= u.size();
    const double n = (double)u.size();
    if (std::fabs(mu - 0.5) > 5 / std::sqrt(12 * n) || std::fabs(var - 1.0 / 12) > 0.01 ||
        std::fabs(gm) > 5 / std::sqrt(n) || std::fabs(gv - 1.0) > 0.05) { printf("FAIL moments %g %g %g %g\n", mu, var, gm, gv); return 1; }
    printf("OK\n");
    return 0;
}
