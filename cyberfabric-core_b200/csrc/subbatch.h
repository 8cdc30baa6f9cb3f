// subbatch.h -- how a pipelined host call cuts a batch into sub-batches (host code; cfbpe.cu:run_host_pipelined uses it,
// tests/simt/sim_harness.cpp exports it to the CPU tests).
#pragma once
#include <stdint.h>

namespace cfbpe {

// Cuts n prompts (byte offsets[0..n], total = offsets[n]) into at most max_chunks runs of whole prompts; cut[k] .. cut[k+1] are
// the prompts of sub-batch k, cut[0] = 0, cut[result] = n; every sub-batch holds at least one prompt.  Sizes ramp up from a small
// first sub-batch (the kernels start early) to `chunk` bytes (big sub-batches keep the kernels efficient) and down again to a
// small last one (little left to download when the kernels end):
//   edge, 2 edge, 4 edge, chunk ... chunk, 4 edge, 2 edge, edge      with edge = chunk / 8
// A prompt longer than its slot simply makes that sub-batch bigger.
inline int plan_sub_batches(const uint64_t* offsets, uint32_t n, uint64_t total, uint64_t chunk, int max_chunks, uint32_t* cut) {
    if (chunk == 0) chunk = 1;
    if ((total + chunk - 1) / chunk + 8 > static_cast<uint64_t>(max_chunks)) chunk = (total + max_chunks - 9) / (max_chunks - 8);
    uint64_t sizes[72];
    int ns = 0;
    {
        uint64_t ramp[8]; int nr = 0;
        for (uint64_t e = chunk / 8 ? chunk / 8 : 1; e < chunk && nr < 3; e *= 2) ramp[nr++] = e;
        uint64_t ramps = 0;
        for (int i = 0; i < nr; ++i) ramps += 2 * ramp[i];
        while (nr && ramps > total) { ramps -= 2 * ramp[nr - 1]; --nr; }
        const uint64_t middle = total - ramps;
        uint64_t n_mid = (middle + chunk - 1) / chunk;
        if (n_mid + 2 * nr > 64) n_mid = 64 - 2 * nr;
        for (int i = 0; i < nr; ++i) sizes[ns++] = ramp[i];
        for (uint64_t i = 0; i < n_mid; ++i) sizes[ns++] = (middle + n_mid - 1) / n_mid;
        for (int i = nr - 1; i >= 0; --i) sizes[ns++] = ramp[i];
        if (!ns) sizes[ns++] = total ? total : 1;
    }
    int nc = 0;
    cut[0] = 0;
    uint32_t p = 0;
    uint64_t target = 0;
    for (int k = 0; k < ns && p < n; ++k) {
        target += sizes[k];
        if (k == ns - 1 || target > total) target = total;
        if (offsets[p] >= target && k < ns - 1) continue;      // a long prompt already covered this slot
        uint32_t lo = p + 1, hi = n;                 // first q > p with offsets[q] >= target (or n)
        while (lo < hi) { const uint32_t m2 = lo + (hi - lo) / 2; if (offsets[m2] >= target) hi = m2; else lo = m2 + 1; }
        p = lo;
        if (nc + 1 == max_chunks) p = n;
        cut[++nc] = p;
    }
    if (p < n) { if (nc && nc == max_chunks) cut[nc] = n; else cut[++nc] = n; }
    return nc;
}

}  // namespace cfbpe
