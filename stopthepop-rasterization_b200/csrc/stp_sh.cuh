// stp_sh.cuh -- spherical-harmonics colour (computeColorFromSH, forward_common.h:20-70), shared by the forward pass
// (preprocess.cu) and by the backward pass (preprocess_bwd.cu), which re-evaluates it on the SH row it has staged anyway
// to obtain the clamp flags: same function, same inputs, hence the same bits as the forward decision -- and no
// dependency of the backward pass on a per-Gaussian `clamped` array that a tile-sharded forward would have to write for
// Gaussians it never renders.
#pragma once
#include "stp_math.cuh"

namespace stp {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
static __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// SH -> RGB (computeColorFromSH, forward_common.h:20-70); sh points at this Gaussian's
// coefficients in shared memory, stride 3 floats per coefficient.
__device__ __forceinline__ void eval_sh(int deg, const float* __restrict__ sh, float dx, float dy, float dz,
                                        float* __restrict__ rgb, uint8_t* __restrict__ clamped3) {
    const float len = fsqrt(ffma(dz, dz, ffma(dx, dx, fmul(dy, dy))));
    const float x = fdiv(dx, len), y = fdiv(dy, len), z = fdiv(dz, len);
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = SH_C0 * sh[c];
    if (deg > 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = r[c] - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                r[c] = r[c] + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                       SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                       SH_C2[4] * (xx - yy) * sh[24 + c];
            if (deg > 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    r[c] = r[c] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                           SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                           SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                           SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                           SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] += 0.5f;
        clamped3[c] = (r[c] < 0.0f) ? 1 : 0;
        rgb[c] = fmaxf(r[c], 0.0f);
    }
}

}  // namespace stp
