// stp_sort.cuh -- warp-level bitonic compare-exchange stage on 64-bit records held in registers (used by the per-tile
// depth sort of binning.cu and by the per-pixel survivor sort of render_ppx.cu).
// A warp keeps a span of 32*E consecutive elements, lane L holding elements L, L+32, ... : compare-exchange
// distances below 32 are warp shuffles, distances 32 .. 16*E are register-to-register.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stp {

template <int E>
__device__ __forceinline__ void reg_stage(uint64_t (&v)[E], int j, int k, int base, int lane) {
    if (j >= 32) {
        const int jj = j >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if ((e & jj) == 0) {
                const bool up = ((base + e * 32 + lane) & k) == 0;
                const uint64_t a = v[e], b = v[e | jj];
                const bool sw = (a > b) == up;
                v[e] = sw ? b : a;
                v[e | jj] = sw ? a : b;
            }
        }
    } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool up = ((base + e * 32 + lane) & k) == 0;
            const uint64_t o = __shfl_xor_sync(0xffffffffu, v[e], j);
            v[e] = ((v[e] < o) == (lower == up)) ? v[e] : o;
        }
    }
}

}  // namespace stp
