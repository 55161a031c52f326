// preprocess_bwd.cu -- per-Gaussian gradient propagation, one fused kernel.
//
// Replaces: computeCov2DCUDA (backward.cu:146-312) + preprocessCUDA<3> backward (backward.cu:384-434)
// with its device functions computeColorFromSH backward (backward.cu:22-141) and computeCov3D
// backward (backward.cu:316-379).  The reference runs two kernels that both re-read the mean, radii
// and (through global memory) dL_dcov3D / dL_dmean3D; here one thread keeps them in registers.
// Gaussians with radii<=0 get zero rows written here (the caller does not pre-clear the outputs).
// Gradients are not integer-decision inputs, so this file uses ordinary float expressions
// (tolerance 1e-5 relative vs. the reference).
#include "stp_kernels.cuh"
#include "stp_sh.cuh"

namespace stp {

namespace {

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 operator*(float s, F3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// The 32 SH rows of a warp (and their gradient rows) are one contiguous span: they are moved between
// global and shared memory with 128-bit, fully coalesced accesses; each lane then works on "its" row
// in shared memory (row stride n_sh+1 floats: conflict-free).  The reference reads and writes the 48
// coefficients per thread with a 192-byte stride (backward.cu:22-141).
__device__ __forceinline__ void preprocess_bwd_one(const PreprocessBwdArgs& a, const Frame& f, int idx, float* __restrict__ row);

template <bool STORE>
__device__ __forceinline__ void move_sh_rows(float* __restrict__ gptr, int rows, int n_sh, uint32_t live, int lane,
                                             float* __restrict__ my_rows, int stride) {
    const int total = rows * n_sh;
    float4* __restrict__ g4 = reinterpret_cast<float4*>(gptr);
    const bool fast = n_sh == 48;
#pragma unroll 4
    for (int i = lane; 4 * i < total; i += 32) {
        const int f0 = 4 * i, f3 = min(f0 + 3, total - 1);
        const int r0 = fast ? (i / 12) : (f0 / n_sh), r3 = fast ? r0 : (f3 / n_sh);
        if (!(((live >> r0) | (live >> r3)) & 1u)) continue;
        int row = r0, col = f0 - r0 * n_sh;
        if (f0 + 3 < total && (fast || r0 == r3)) {
            if constexpr (STORE) {
                float4 v;
                v.x = my_rows[row * stride + col];
                v.y = my_rows[row * stride + col + 1];
                v.z = my_rows[row * stride + col + 2];
                v.w = my_rows[row * stride + col + 3];
                g4[i] = v;
            } else {
                const float4 v = __ldg(g4 + i);
                my_rows[row * stride + col] = v.x;
                my_rows[row * stride + col + 1] = v.y;
                my_rows[row * stride + col + 2] = v.z;
                my_rows[row * stride + col + 3] = v.w;
            }
        } else {
            for (int k = 0; k < 4 && f0 + k < total; ++k) {
                if ((live >> row) & 1u) {
                    if constexpr (STORE) gptr[f0 + k] = my_rows[row * stride + col];
                    else my_rows[row * stride + col] = __ldg(gptr + f0 + k);
                }
                if (++col == n_sh) {
                    col = 0;
                    ++row;
                }
            }
        }
    }
}

#ifndef STP_PREBWD_MINB
#define STP_PREBWD_MINB 4  // 64 registers, 4 CTAs/SM: A/B on B200 (C2: 0.23 -> 0.16 ms, C5: 2.14 -> 1.48 ms)
#endif
__global__ void __launch_bounds__(256, STP_PREBWD_MINB)
preprocess_bwd_kernel(PreprocessBwdArgs a, Frame f) {
    extern __shared__ float s_rows[];  // [8 warps][32 rows][3M+1]
    // Gaussians [a.first, a.P_end): the launch can cover a sub-range so that a data-parallel caller can start the
    // all-reduce of one range of gradient rows while the next range is still being computed (a.first % 256 == 0)
    const int idx = a.first + blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = idx < a.P_end;
    const bool alive = valid && a.radii[idx] > 0;
    const uint32_t live = __ballot_sync(0xffffffffu, alive);
    const int n_sh = a.M * 3, sh_stride = n_sh + 1;
    const int warp_base = a.first + blockIdx.x * blockDim.x + warp * 32;
    const int rows = min(32, a.P_end - warp_base);
    const bool has_sh = a.shs != nullptr && n_sh > 0;
    // Every row of every output (dL_dmean2D / dL_dcolor / dL_dopacity included) is written here: culled Gaussians get
    // zeros, so the caller does not have to clear these buffers (the reference memsets 256 B per Gaussian per
    // iteration for them, rasterize_points.cu:178-186).
    if (valid && !alive) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            a.dL_dmean3D[3 * idx + k] = 0.f;
            a.dL_dmean2D[3 * idx + k] = 0.f;
            a.dL_dcolor[3 * idx + k] = 0.f;
        }
        a.dL_dopacity[idx] = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * idx + k] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) a.dL_dscale[3 * idx + k] = 0.f;
        reinterpret_cast<float4*>(a.dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool rows_aligned = has_sh && ((reinterpret_cast<uintptr_t>(a.shs + (size_t)warp_base * n_sh) |
                                          reinterpret_cast<uintptr_t>(a.dL_dsh + (size_t)warp_base * n_sh)) & 15u) == 0;
    if (live == 0) {
        if (has_sh && rows > 0) {  // whole warp culled: clear its 32 gradient rows with coalesced stores
            float* dst = a.dL_dsh + (size_t)warp_base * n_sh;
            const int total = rows * n_sh;
            if (rows_aligned && (total & 3) == 0) {
                for (int i = lane; 4 * i < total; i += 32) reinterpret_cast<float4*>(dst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int i = lane; i < total; i += 32) dst[i] = 0.f;
            }
        }
        return;
    }
    float* const my_rows = s_rows + (size_t)warp * 32 * sh_stride;
    const bool sh_staged = rows_aligned;
    if (sh_staged) {
        if (n_sh == 48)
            move_sh_rows_48<false>(const_cast<float*>(a.shs) + (size_t)warp_base * n_sh, rows, live, lane, my_rows);
        else
            move_sh_rows<false>(const_cast<float*>(a.shs) + (size_t)warp_base * n_sh, rows, n_sh, live, lane, my_rows, sh_stride);
        __syncwarp();
    }
    if (alive) {
        preprocess_bwd_one(a, f, idx, sh_staged ? my_rows + lane * sh_stride : nullptr);
    } else if (valid && has_sh) {
        if (sh_staged) {
            for (int k = 0; k < n_sh; ++k) my_rows[lane * sh_stride + k] = 0.f;
        } else {
            for (int k = 0; k < n_sh; ++k) a.dL_dsh[(size_t)idx * n_sh + k] = 0.f;
        }
    }
    if (sh_staged) {
        __syncwarp();
        const uint32_t all_rows = rows >= 32 ? 0xffffffffu : ((1u << rows) - 1u);
        if (n_sh == 48)
            move_sh_rows_48<true>(a.dL_dsh + (size_t)warp_base * n_sh, rows, all_rows, lane, my_rows);
        else
            move_sh_rows<true>(a.dL_dsh + (size_t)warp_base * n_sh, rows, n_sh, all_rows, lane, my_rows, sh_stride);
    }
}

// everything of one Gaussian; `row` = its SH coefficients in shared memory (overwritten with dL_dsh), or nullptr to
// read/write global memory directly
__device__ __forceinline__ void preprocess_bwd_one(const PreprocessBwdArgs& a, const Frame& f, int idx, float* __restrict__ row) {
    const float* __restrict__ vm = f.viewmatrix;
    const float* __restrict__ proj = f.projmatrix;

    const F3 mean = {a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]};
    float c3[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) c3[k] = a.cov3D[6 * idx + k];

    // ---------------- conic -> cov2D -> cov3D / mean (computeCov2DCUDA) ----------------
    // packed screen-space gradients of this Gaussian (stp_kernels.cuh: kGradAccum layout)
    const float4 acc0 = reinterpret_cast<const float4*>(a.grad_accum)[idx];                       // plane A
    const float4 acc1 = reinterpret_cast<const float4*>(a.grad_accum)[(size_t)a.P + idx];          // plane B
    const float acc_cb = a.grad_accum[8 * (size_t)a.P + idx];                                     // plane C
    const float dcon_x = acc0.x, dcon_y = acc0.y, dcon_z = acc0.z;
    a.dL_dmean2D[3 * idx] = acc1.x;
    a.dL_dmean2D[3 * idx + 1] = acc1.y;
    a.dL_dmean2D[3 * idx + 2] = 0.f;
    a.dL_dcolor[3 * idx] = acc1.z;
    a.dL_dcolor[3 * idx + 1] = acc1.w;
    a.dL_dcolor[3 * idx + 2] = acc_cb;
    a.dL_dopacity[idx] = acc0.w;
    F3 t = {vm[0] * mean.x + vm[4] * mean.y + vm[8] * mean.z + vm[12], vm[1] * mean.x + vm[5] * mean.y + vm[9] * mean.z + vm[13],
            vm[2] * mean.x + vm[6] * mean.y + vm[10] * mean.z + vm[14]};
    const float limx = 1.3f * f.tan_fovx, limy = 1.3f * f.tan_fovy;
    const float txtz = t.x / t.z, tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float h_x = f.focal_x, h_y = f.focal_y;

    // J (column-major like glm): J[0]=(J00,0,J02) J[1]=(0,J11,J12)
    const float J00 = h_x / t.z, J02 = -(h_x * t.x) / (t.z * t.z);
    const float J11 = h_y / t.z, J12 = -(h_y * t.y) / (t.z * t.z);
    // W[c][r] = vm[4r + c];  T = W*J, T[c][r]
    float T0[3], T1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T0[r] = vm[4 * r + 0] * J00 + vm[4 * r + 2] * J02;
        T1[r] = vm[4 * r + 1] * J11 + vm[4 * r + 2] * J12;
    }
    // Vrk symmetric
    const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float VT0[3], VT1[3];  // Vrk * T0, Vrk * T1
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        VT0[r] = V[r][0] * T0[0] + V[r][1] * T0[1] + V[r][2] * T0[2];
        VT1[r] = V[r][0] * T1[0] + V[r][1] * T1[1] + V[r][2] * T1[2];
    }
    float c_xx = T0[0] * VT0[0] + T0[1] * VT0[1] + T0[2] * VT0[2];
    const float c_xy = T0[0] * VT1[0] + T0[1] * VT1[1] + T0[2] * VT1[2];
    float c_yy = T1[0] * VT1[0] + T1[1] * VT1[1] + T1[2] * VT1[2];
    const float det_cov_orig = c_xx * c_yy - c_xy * c_xy;
    constexpr float h_var = 0.3f;
    c_xx += h_var;
    c_yy += h_var;

    float dL_dc_xx = 0.f, dL_dc_xy = 0.f, dL_dc_yy = 0.f;
    if (a.proper_ewa_scaling) {  // backward.cu:214-238
        const float det_plus = c_xx * c_yy - c_xy * c_xy;
        const float ratio = det_cov_orig / det_plus;
        const float h_scaling = sqrtf(fmaxf(0.000025f, ratio));
        const float dL_dop = acc0.w;
        const float d_h = dL_dop * a.opacities[idx];
        a.dL_dopacity[idx] = dL_dop * h_scaling;
        const float d_inside_root = (ratio <= 0.000025f) ? 0.f : d_h / (2.f * h_scaling);
        const float x = c_xx, y = c_yy, z = c_xy, w = h_var;
        const float q = w * w + w * (x + y) + x * y - z * z;
        const float denom_f = d_inside_root / (q * q);
        dL_dc_xx = w * (w * y + y * y + z * z) * denom_f;
        dL_dc_yy = w * (w * x + x * x + z * z) * denom_f;
        dL_dc_xy = -2.f * w * z * (w + x + y) * denom_f;
    }

    const float denom = c_xx * c_yy - c_xy * c_xy;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcov[6];
    if (denom2inv != 0.f) {
        dL_dc_xx += denom2inv * (-c_yy * c_yy * dcon_x + 2.f * c_xy * c_yy * dcon_y + (denom - c_xx * c_yy) * dcon_z);
        dL_dc_yy += denom2inv * (-c_xx * c_xx * dcon_z + 2.f * c_xx * c_xy * dcon_y + (denom - c_xx * c_yy) * dcon_x);
        dL_dc_xy += denom2inv * 2.f * (c_xy * c_yy * dcon_x - (denom + 2.f * c_xy * c_xy) * dcon_y + c_xx * c_xy * dcon_z);
        dcov[0] = T0[0] * T0[0] * dL_dc_xx + T0[0] * T1[0] * dL_dc_xy + T1[0] * T1[0] * dL_dc_yy;
        dcov[3] = T0[1] * T0[1] * dL_dc_xx + T0[1] * T1[1] * dL_dc_xy + T1[1] * T1[1] * dL_dc_yy;
        dcov[5] = T0[2] * T0[2] * dL_dc_xx + T0[2] * T1[2] * dL_dc_xy + T1[2] * T1[2] * dL_dc_yy;
        dcov[1] = 2.f * T0[0] * T0[1] * dL_dc_xx + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_dc_xy + 2.f * T1[0] * T1[1] * dL_dc_yy;
        dcov[2] = 2.f * T0[0] * T0[2] * dL_dc_xx + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_dc_xy + 2.f * T1[0] * T1[2] * dL_dc_yy;
        dcov[4] = 2.f * T0[2] * T0[1] * dL_dc_xx + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_dc_xy + 2.f * T1[1] * T1[2] * dL_dc_yy;
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) dcov[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * idx + k] = dcov[k];

    // dL/dT (upper 2x3), backward.cu:270-283
    const float dT00 = 2.f * VT0[0] * dL_dc_xx + VT1[0] * dL_dc_xy;
    const float dT01 = 2.f * VT0[1] * dL_dc_xx + VT1[1] * dL_dc_xy;
    const float dT02 = 2.f * VT0[2] * dL_dc_xx + VT1[2] * dL_dc_xy;
    const float dT10 = 2.f * VT1[0] * dL_dc_yy + VT0[0] * dL_dc_xy;
    const float dT11 = 2.f * VT1[1] * dL_dc_yy + VT0[1] * dL_dc_xy;
    const float dT12 = 2.f * VT1[2] * dL_dc_yy + VT0[2] * dL_dc_xy;
    // dL/dJ, W[c][r] = vm[4r+c]
    const float dJ00 = vm[0] * dT00 + vm[4] * dT01 + vm[8] * dT02;
    const float dJ02 = vm[2] * dT00 + vm[6] * dT01 + vm[10] * dT02;
    const float dJ11 = vm[1] * dT10 + vm[5] * dT11 + vm[9] * dT12;
    const float dJ12 = vm[2] * dT10 + vm[6] * dT11 + vm[10] * dT12;
    const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dL_dtx = x_grad_mul * -h_x * tz2 * dJ02;
    const float dL_dty = y_grad_mul * -h_y * tz2 * dJ12;
    const float dL_dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2.f * h_x * t.x) * tz3 * dJ02 + (2.f * h_y * t.y) * tz3 * dJ12;
    // transformVec4x3Transpose
    F3 dmean = {vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz, vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
                vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz};

    // ---------------- mean2D -> mean3D (preprocessCUDA bwd, backward.cu:408-421) ----------------
    {
        const float hw = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        const float gx = acc1.x, gy = acc1.y;
        dmean.x += (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
        dmean.y += (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
        dmean.z += (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
    }

    // ---------------- colour -> SH and view direction (computeColorFromSH bwd) ----------------
    if (a.shs != nullptr) {
        const F3 dir_orig = {mean.x - f.cam_pos[0], mean.y - f.cam_pos[1], mean.z - f.cam_pos[2]};
        const float len = sqrtf(dot(dir_orig, dir_orig));
        const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
        // no __restrict__: in the staged path both point at the same shared-memory row, which is rewritten in place
        const float* shp = row ? row : a.shs + (size_t)idx * a.M * 3;
        float* dsh = row ? row : a.dL_dsh + (size_t)idx * a.M * 3;
        const int D = a.D;
        auto sh = [&](int k) { return F3{shp[3 * k], shp[3 * k + 1], shp[3 * k + 2]}; };
        F3 dRGB = {acc1.z, acc1.w, acc_cb};
        {
            // clamp flags of the forward pass (computeColorFromSH sets them where a channel went negative): re-evaluated
            // with the forward's own function on the same row and direction, before the row is overwritten below
            float rgb_unused[3];
            uint8_t cl[3];
            eval_sh(D, shp, dir_orig.x, dir_orig.y, dir_orig.z, rgb_unused, cl);
            dRGB.x *= cl[0] ? 0.f : 1.f;
            dRGB.y *= cl[1] ? 0.f : 1.f;
            dRGB.z *= cl[2] ? 0.f : 1.f;
        }
        // weights of dL_dsh[k] = w[k] * dRGB; written after every coefficient has been read (the row is
        // updated in place when it lives in shared memory)
        float wk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) wk[k] = 0.f;
        auto put = [&](int k, float w) { wk[k] = w; };
        F3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
        put(0, SH_C0);
        if (D > 0) {
            put(1, -SH_C1 * y);
            put(2, SH_C1 * z);
            put(3, -SH_C1 * x);
            dx = -SH_C1 * sh(3);
            dy = -SH_C1 * sh(1);
            dz = SH_C1 * sh(2);
            if (D > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                put(4, SH_C2[0] * xy);
                put(5, SH_C2[1] * yz);
                put(6, SH_C2[2] * (2.f * zz - xx - yy));
                put(7, SH_C2[3] * xz);
                put(8, SH_C2[4] * (xx - yy));
                dx = dx + (SH_C2[0] * y) * sh(4) + (SH_C2[2] * 2.f * -x) * sh(6) + (SH_C2[3] * z) * sh(7) + (SH_C2[4] * 2.f * x) * sh(8);
                dy = dy + (SH_C2[0] * x) * sh(4) + (SH_C2[1] * z) * sh(5) + (SH_C2[2] * 2.f * -y) * sh(6) + (SH_C2[4] * 2.f * -y) * sh(8);
                dz = dz + (SH_C2[1] * y) * sh(5) + (SH_C2[2] * 2.f * 2.f * z) * sh(6) + (SH_C2[3] * x) * sh(7);
                if (D > 2) {
                    put(9, SH_C3[0] * y * (3.f * xx - yy));
                    put(10, SH_C3[1] * xy * z);
                    put(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                    put(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                    put(13, SH_C3[4] * x * (4.f * zz - xx - yy));
                    put(14, SH_C3[5] * z * (xx - yy));
                    put(15, SH_C3[6] * x * (xx - 3.f * yy));
                    dx = dx + (SH_C3[0] * 3.f * 2.f * xy) * sh(9) + (SH_C3[1] * yz) * sh(10) + (SH_C3[2] * -2.f * xy) * sh(11) +
                         (SH_C3[3] * -3.f * 2.f * xz) * sh(12) + (SH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * sh(13) +
                         (SH_C3[5] * 2.f * xz) * sh(14) + (SH_C3[6] * 3.f * (xx - yy)) * sh(15);
                    dy = dy + (SH_C3[0] * 3.f * (xx - yy)) * sh(9) + (SH_C3[1] * xz) * sh(10) +
                         (SH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * sh(11) + (SH_C3[3] * -3.f * 2.f * yz) * sh(12) +
                         (SH_C3[4] * -2.f * xy) * sh(13) + (SH_C3[5] * -2.f * yz) * sh(14) + (SH_C3[6] * -3.f * 2.f * xy) * sh(15);
                    dz = dz + (SH_C3[1] * xy) * sh(10) + (SH_C3[2] * 4.f * 2.f * yz) * sh(11) +
                         (SH_C3[3] * 3.f * (2.f * zz - xx - yy)) * sh(12) + (SH_C3[4] * 4.f * 2.f * xz) * sh(13) +
                         (SH_C3[5] * (xx - yy)) * sh(14);
                }
            }
        }
        {
            // every one of the M coefficients is written on both paths (wk is zero beyond the active degree): dL_dsh is a
            // pure output that the caller does not zero-fill
            const int ncoef = a.M;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < ncoef) {
                    dsh[3 * k] = wk[k] * dRGB.x;
                    dsh[3 * k + 1] = wk[k] * dRGB.y;
                    dsh[3 * k + 2] = wk[k] * dRGB.z;
                }
            }
        }
        const F3 ddir = {dot(dx, dRGB), dot(dy, dRGB), dot(dz, dRGB)};
        // dnormvdv, auxiliary.h:181-191
        const F3 v = dir_orig;
        const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
        const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dmean.x += ((+sum2 - v.x * v.x) * ddir.x - v.y * v.x * ddir.y - v.z * v.x * ddir.z) * invsum32;
        dmean.y += (-v.x * v.y * ddir.x + (sum2 - v.y * v.y) * ddir.y - v.z * v.y * ddir.z) * invsum32;
        dmean.z += (-v.x * v.z * ddir.x - v.y * v.z * ddir.y + (sum2 - v.z * v.z) * ddir.z) * invsum32;
    }
    a.dL_dmean3D[3 * idx] = dmean.x;
    a.dL_dmean3D[3 * idx + 1] = dmean.y;
    a.dL_dmean3D[3 * idx + 2] = dmean.z;

    // ---------------- cov3D -> scale / rotation (computeCov3D bwd, backward.cu:316-379) ----------------
    if (a.scales != nullptr) {
        const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        // R[c][r] (glm column-major)
        const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                               {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                               {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {a.scale_modifier * a.scales[3 * idx], a.scale_modifier * a.scales[3 * idx + 1],
                            a.scale_modifier * a.scales[3 * idx + 2]};
        // M = S*R : M[c][r] = s[r]*R[c][r]
        float M[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) M[c][rr] = s[rr] * R[c][rr];
        // dL_dSigma (symmetric), column-major
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        // dL_dM = 2 * M * dL_dSigma : (A*B)[c][r] = sum_k A[k][r]*B[c][k]
        float dM[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
                dM[c][rr] = 2.0f * M[0][rr] * dS[c][0] + 2.0f * M[1][rr] * dS[c][1] + 2.0f * M[2][rr] * dS[c][2];
        // Rt[c][r] = R[r][c]; dMt[c][r] = dM[r][c]
        float dscale[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) dscale[k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
        a.dL_dscale[3 * idx] = dscale[0];
        a.dL_dscale[3 * idx + 1] = dscale[1];
        a.dL_dscale[3 * idx + 2] = dscale[2];
        // dMt[k] *= s[k]   (dMt[k][j] = dM[j][k])
        float dMt[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) dMt[k][j] = dM[j][k] * s[k];
        float4 dq;
        dq.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        dq.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
               4 * x * (dMt[2][2] + dMt[1][1]);
        dq.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
               4 * y * (dMt[2][2] + dMt[0][0]);
        dq.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
               4 * z * (dMt[1][1] + dMt[0][0]);
        reinterpret_cast<float4*>(a.dL_drot)[idx] = dq;
    } else {  // cov3D_precomp: no scale / rotation gradient (the outputs are pure outputs, so they are still written)
        a.dL_dscale[3 * idx] = 0.f;
        a.dL_dscale[3 * idx + 1] = 0.f;
        a.dL_dscale[3 * idx + 2] = 0.f;
        reinterpret_cast<float4*>(a.dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace

cudaError_t launch_preprocess_bwd(const PreprocessBwdArgs& a, const Frame& f, cudaStream_t stream) {
    const size_t smem = (a.shs != nullptr && a.M > 0) ? sizeof(float) * 8 * 32 * (a.M * 3 + 1) : 0;
    cudaFuncSetAttribute(preprocess_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int count = a.P_end - a.first;
    if (count <= 0) return cudaSuccess;
    preprocess_bwd_kernel<<<(count + 255) / 256, 256, smem, stream>>>(a, f);
    return cudaGetLastError();
}

}  // namespace stp
