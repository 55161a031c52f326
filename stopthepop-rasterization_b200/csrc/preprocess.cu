// preprocess.cu -- per-Gaussian projection, culling, tile counting (+ the per-tile instance histogram that sizes
// the tile buckets of binning.cu) and SH->RGB in ONE kernel.
//
// Replaces: preprocessCUDA<3,TBC,LB> (forward.cu:68-229) + checkFrustum (rasterizer_impl.cu:113-128); the
// reference's cub::DeviceScan::InclusiveSum over tiles_touched (rasterizer_impl.cu:313) has no counterpart: instance
// slots are claimed per tile, the only prefix sum left is over the tiles (tile_scan_kernel, binning.cu).
//
// B200 design notes
//   * one thread per Gaussian, 256-thread CTAs, grid sized by P (memory-bound: ~236 B in, ~90-140 B out
//     per Gaussian); no inter-CTA dependency of any kind.
//   * SH coefficients (192 of the 236 input bytes) are only read for survivors and are pulled
//     warp-cooperatively: the 32 lanes stream one survivor's contiguous 12*M bytes with coalesced
//     loads into shared memory (row stride M*3+1 -> conflict-free), instead of 48 strided scalar
//     loads per thread.
//   * the exact-tile-count loop of tile_based_culling is always load balanced: the first
//     kSeqTiles tiles of a rectangle are tested by the owning thread, the remainder by the whole warp.
//     Results do not depend on the schedule (settings.load_balancing is a no-op by construction).
#include "stp_kernels.cuh"
#include "stp_sh.cuh"

namespace stp {

namespace {

constexpr int kSeqTiles = 8;  // tiles of a rectangle tested sequentially by the owner thread

// does tile (tx,ty) pass the exact contribution test? (computeTilebasedCullingTileCount,
// stopthepop_common.cuh:176-262; same arithmetic in duplicateWithKeys_extended :419-452)
__device__ __forceinline__ bool tile_contributes(float A, float B, float C, float2 xy, float thr, int tx, int ty) {
    float mx, my;
    const float p = max_contrib_power<15, 15>(A, B, C, xy.x, xy.y, (float)(tx * 16), (float)(ty * 16),
                                              (float)(tx * 16 + 15), (float)(ty * 16 + 15), mx, my);
    return p <= thr;
}

// The 32 coefficient rows of a warp are one contiguous span of rows*n_sh floats: stream it with 128-bit
// loads (several independent loads in flight per lane) into the warp's shared-memory rows, skipping
// quads whose rows were all culled.  NSH != 0 = compile-time row length (48 at SH degree 3).
template <int NSH>
__device__ __forceinline__ void stage_sh_rows(const float* __restrict__ wsrc, int rows, int n_sh_rt, uint32_t surv, int lane,
                                              float* __restrict__ my_rows, int stride) {
    const int n_sh = NSH ? NSH : n_sh_rt;
    const int total = rows * n_sh;  // floats
    const float4* __restrict__ src4 = reinterpret_cast<const float4*>(wsrc);
#pragma unroll 4
    for (int i = lane; 4 * i < total; i += 32) {
        const int f0 = 4 * i, f3 = min(f0 + 3, total - 1);
        const int r0 = f0 / n_sh, r3 = f3 / n_sh;
        if (((surv >> r0) | (surv >> r3)) & 1u) {
            float4 v;
            if (f0 + 3 < total) {
                v = __ldg(src4 + i);
            } else {
                v.x = __ldg(wsrc + f0);
                v.y = (f0 + 1 < total) ? __ldg(wsrc + f0 + 1) : 0.f;
                v.z = (f0 + 2 < total) ? __ldg(wsrc + f0 + 2) : 0.f;
                v.w = 0.f;
            }
            const float e[4] = {v.x, v.y, v.z, v.w};
            int row = r0, col = f0 - r0 * n_sh;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (f0 + k < total) my_rows[row * stride + col] = e[k];
                if (++col == n_sh) {
                    col = 0;
                    ++row;
                }
            }
        }
    }
}

}  // namespace

template <bool TBC, bool BANDED>
#ifndef STP_PRE_MINB
#define STP_PRE_MINB 4  // 64 registers, 4 CTAs/SM: A/B on B200 (C5: 1.83 -> 1.47 ms)
#endif
__global__ void __launch_bounds__(kPreprocessThreads, STP_PRE_MINB)
preprocess_kernel(PreprocessArgs a, Frame f, GeometryState g, uint32_t* __restrict__ tile_count) {
    extern __shared__ float s_sh[];  // [8 warps][32 rows][sh_stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bid = blockIdx.x;
    const int idx = bid * kPreprocessThreads + tid;
    const bool valid = idx < a.P;

    bool alive = valid;
    uint32_t tiles = 0;
    float x = 0.f, y = 0.f, z = 0.f;
    Vec3 pv{0.f, 0.f, 1.f};
    float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 mean2D = make_float2(0.f, 0.f), rect_ext = make_float2(0.f, 0.f);
    float radius = 0.f, thr = 0.f;
    TileRect rc{0, 0, 0, 0};

    // all per-Gaussian inputs are requested up front (independent loads in flight) although culled
    // Gaussians will not use them: the kernel is latency-, not bandwidth-limited
    float4 q_in = make_float4(0.f, 0.f, 0.f, 0.f);
    float s_in[3] = {0.f, 0.f, 0.f};
    float opacity_in = 0.f;
    if (valid) {
        x = __ldg(a.means3D + 3 * idx + 0);
        y = __ldg(a.means3D + 3 * idx + 1);
        z = __ldg(a.means3D + 3 * idx + 2);
        opacity_in = __ldg(a.opacities + idx);
        if (a.cov3D_precomp == nullptr) {
            q_in = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
            s_in[0] = __ldg(a.scales + 3 * idx + 0);
            s_in[1] = __ldg(a.scales + 3 * idx + 1);
            s_in[2] = __ldg(a.scales + 3 * idx + 2);
        }
        pv = view_transform(f.viewmatrix, x, y, z);
        if (pv.z <= kNearPlane) {  // in_frustum, auxiliary.h:223
            alive = false;
            if (a.prefiltered) atomicOr(&g.counters[2], 1u);
        }
    }

    float cov6[6];
    if (alive) {
        if (a.cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) cov6[k] = a.cov3D_precomp[6 * idx + k];
        } else {
            const float4 q = q_in;
            const Rot3 R = quat_to_rot(q.x, q.y, q.z, q.w);
            const float sx = s_in[0], sy = s_in[1], sz = s_in[2];
            gram_scaled_rot(R, fmul(a.scale_modifier, sx), fmul(a.scale_modifier, sy), fmul(a.scale_modifier, sz), cov6);
#pragma unroll
            for (int k = 0; k < 6; ++k) g.cov3D[6 * idx + k] = cov6[k];
        }
        const Vec3 cov = project_cov2d(pv, f.focal_x, f.focal_y, f.tan_fovx, f.tan_fovy, cov6, f.viewmatrix);
        // dilateCov2D + computeConicOpacity, forward_common.h:108-144
        const float ca = fadd(cov.x, 0.3f), cc = fadd(cov.z, 0.3f), cb = cov.y;
        const float bb = fmul(cb, cb);
        const float det = ffma(ca, cc, -bb);
        float scaling = 1.0f;
        if (a.proper_ewa_scaling) {
            const float det_orig = ffma(cov.x, cov.z, -bb);
            scaling = fsqrt(fmaxf(0.000025f, fdiv(det_orig, det)));
        }
        if (det == 0.0f) {
            alive = false;
        } else {
            const float det_inv = fdiv(1.0f, det);
            co.x = fmul(cc, det_inv);
            co.y = fmul(cb, -det_inv);
            co.z = fmul(ca, det_inv);
            co.w = fmul(opacity_in, scaling);
            if (co.w < kAlphaThreshold) alive = false;
        }
        if (alive) {
            thr = logf(fdiv(co.w, kAlphaThreshold));
            float extent = 3.33f;
            if (a.tight_opacity_bounding) extent = (float)fmin(3.33, (double)fsqrt(fadd(thr, thr)));
            const float mid = fmul(0.5f, fadd(ca, cc));
            const float lambda = fadd(mid, fsqrt(fmaxf(0.01f, ffma(mid, mid, -det))));
            radius = fmul(extent, fsqrt(lambda));
            if (radius <= 0.0f) alive = false;
            if (alive) {
                mean2D = project_mean2d(f.projmatrix, x, y, z, f.W, f.H);
                rect_ext.x = fminf(a.rect_bounding ? fmul(extent, fsqrt(ca)) : radius, radius);
                rect_ext.y = fminf(a.rect_bounding ? fmul(extent, fsqrt(cc)) : radius, radius);
                // visibility (radii, the geometry state) never depends on a tile band: every rank of a tile-sharded frame
                // holds the state of every visible Gaussian, so that the backward pass can be finished anywhere once
                // the screen-space gradients are summed (stp_sharding.py).  Only the binning below is clamped to the band.
                rc = tile_rect(mean2D, rect_ext, f.grid_x, f.grid_y, 0, f.grid_y);
                tiles = (uint32_t)((rc.x1 - rc.x0) * (rc.y1 - rc.y0));
                if (tiles == 0) alive = false;
            }
        }
    }

    {
        // per-tile instance histogram (sizes the tile buckets of binning.cu) and, with tile_based_culling, the
        // exact tile count: the owner thread visits the first kSeqTiles tiles, the warp shares the rest.
        // Tile band (multi-GPU): only tiles of rows [row0,row1) are binned.  Without culling the walk is clamped to the
        // band; with culling the whole rectangle is walked, because "some tile contributes" decides visibility.
        // (BANDED = false: whole image, every tile of the rectangle is binned -- the single-GPU instantiation carries none
        // of the band bookkeeping)
        const int by0 = BANDED ? min(f.row1, max(f.row0, rc.y0)) : rc.y0, by1 = BANDED ? min(f.row1, max(f.row0, rc.y1)) : rc.y1;
        if constexpr (!TBC && BANDED) {
            rc.y0 = by0;
            rc.y1 = by1;
        }
        const int rect_tiles = alive ? (rc.x1 - rc.x0) * (rc.y1 - rc.y0) : 0;
        const int rw = max(rc.x1 - rc.x0, 1);
        int count = 0, count_band = 0;
        for (int t = 0, tx = rc.x0, ty = rc.y0; t < min(rect_tiles, kSeqTiles); ++t) {
            if (!TBC || tile_contributes(co.x, co.y, co.z, mean2D, thr, tx, ty)) {
                ++count;
                if (!TBC || !BANDED || (ty >= by0 && ty < by1)) {
                    ++count_band;
                    atomicAdd(tile_count + ty * f.grid_x + tx, 1u);
                }
            }
            if (++tx == rc.x1) {
                tx = rc.x0;
                ++ty;
            }
        }
        uint32_t big = __ballot_sync(0xffffffffu, rect_tiles > kSeqTiles);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            float A = 0.f, B = 0.f, C = 0.f, th = 0.f;
            float2 m = make_float2(0.f, 0.f);
            if constexpr (TBC) {
                A = __shfl_sync(0xffffffffu, co.x, src);
                B = __shfl_sync(0xffffffffu, co.y, src);
                C = __shfl_sync(0xffffffffu, co.z, src);
                m = make_float2(__shfl_sync(0xffffffffu, mean2D.x, src), __shfl_sync(0xffffffffu, mean2D.y, src));
                th = __shfl_sync(0xffffffffu, thr, src);
            }
            const int x0 = __shfl_sync(0xffffffffu, rc.x0, src), y0 = __shfl_sync(0xffffffffu, rc.y0, src);
            const int w = __shfl_sync(0xffffffffu, rw, src), n = __shfl_sync(0xffffffffu, rect_tiles, src);
            const int sy0 = __shfl_sync(0xffffffffu, by0, src), sy1 = __shfl_sync(0xffffffffu, by1, src);
            int c = 0, cb = 0;
            for (int t = kSeqTiles + lane; t < n; t += 32) {
                const int tx = x0 + t % w, ty = y0 + t / w;
                if (!TBC || tile_contributes(A, B, C, m, th, tx, ty)) {
                    ++c;
                    if (!TBC || !BANDED || (ty >= sy0 && ty < sy1)) {
                        ++cb;
                        atomicAdd(tile_count + ty * f.grid_x + tx, 1u);
                    }
                }
            }
            if constexpr (TBC) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    if constexpr (BANDED) cb += __shfl_xor_sync(0xffffffffu, cb, o);
                }
                if constexpr (!BANDED) cb = c;
                if (lane == src) {
                    count += c;
                    count_band += cb;
                }
            }
        }
        if (alive) {
            if (TBC && count == 0) alive = false;  // no tile of the image sees it
            tiles = TBC ? (uint32_t)count_band : (uint32_t)rect_tiles;  // instances this rank bins
        }
    }
    if (!alive) tiles = 0;

    // ---- survivors: colour, inverse covariance, stores --------------------------------------
    if (a.colors_precomp == nullptr && a.M > 0) {
        const int stride = a.M * 3 + 1;
        float* my_rows = s_sh + (size_t)warp * 32 * stride;
        // tile band: a Gaussian without an instance in this rank's band is never blended here -- its SH row (81 % of the
        // bytes this kernel reads) is not fetched and its colour not evaluated
        const bool coloured = alive && (!BANDED || tiles != 0u);
        uint32_t surv = __ballot_sync(0xffffffffu, coloured);
        const int warp_base = (int)bid * kPreprocessThreads + warp * 32;
        const int n_sh = a.M * 3;
        const float* __restrict__ wsrc = a.shs + (size_t)warp_base * n_sh;
        const int rows = min(32, a.P - warp_base);
        if (rows > 0 && (reinterpret_cast<uintptr_t>(wsrc) & 15u) == 0 && surv != 0) {
            if (n_sh == 48)
                move_sh_rows_48<false>(const_cast<float*>(wsrc), rows, surv, lane, my_rows);
            else
                stage_sh_rows<0>(wsrc, rows, n_sh, surv, lane, my_rows, stride);
        } else {
            while (surv) {
                const int r = __ffs(surv) - 1;
                surv &= surv - 1;
                const float* __restrict__ src = a.shs + (size_t)(warp_base + r) * n_sh;
                for (int k = lane; k < n_sh; k += 32) my_rows[r * stride + k] = __ldg(src + k);
            }
        }
        __syncwarp();
        if (coloured) {
            float rgb[3];
            uint8_t cl[3];
            eval_sh(a.D, my_rows + lane * stride, fsub(x, f.cam_pos[0]), fsub(y, f.cam_pos[1]), fsub(z, f.cam_pos[2]),
                    rgb, cl);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                g.rgb[3 * idx + c] = rgb[c];
                g.clamped[3 * idx + c] = cl[c];
            }
        }
    }

    if (alive) {
        const float vx = fsub(f.cam_pos[0], x), vy = fsub(f.cam_pos[1], y), vz = fsub(f.cam_pos[2], z);
        if (g.cov3D_inv != nullptr) {
            // computeInvCov3D + packing, stopthepop_common.cuh:13-41, forward.cu:208-220
            const float4 q = q_in;
            const Rot3 R = quat_to_rot(q.x, q.y, q.z, q.w);
            const float sx = s_in[0], sy = s_in[1], sz = s_in[2];
            float ic[6];
            gram_scaled_rot(R, fdiv(1.0f, fmul(a.scale_modifier, fmaxf(1e-3f, sx))),
                            fdiv(1.0f, fmul(a.scale_modifier, fmaxf(1e-3f, sy))),
                            fdiv(1.0f, fmul(a.scale_modifier, fmaxf(1e-3f, sz))), ic);
            const float ux = ffma(-ic[2], vz, ffma(-ic[1], vy, -fmul(ic[0], vx)));
            const float uy = ffma(-ic[4], vz, ffma(-ic[3], vy, -fmul(ic[1], vx)));
            const float uz = ffma(-ic[5], vz, ffma(-ic[4], vy, -fmul(ic[2], vx)));
            g.cov3D_inv[3 * idx + 0] = make_float4(ic[0], ic[1], ic[2], 0.f);
            g.cov3D_inv[3 * idx + 1] = make_float4(ic[3], ic[4], ic[5], 0.f);
            g.cov3D_inv[3 * idx + 2] = make_float4(ux, uy, uz, 0.f);
        }
        g.depths[idx] = (a.sort_order == 0) ? pv.z : fsqrt(ffma(vz, vz, ffma(vx, vx, fmul(vy, vy))));
        g.rects2D[idx] = rect_ext;
        g.means2D[idx] = mean2D;
        g.conic_opacity[idx] = co;
    }
    if (valid) {
        a.radii[idx] = alive ? (int)ceilf(radius) : 0;
        g.tiles_touched[idx] = tiles;
    }

}

// markVisible (rasterizer_impl.cu:113-128): the same near-plane test, nothing else.
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ vm,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const Vec3 pv = view_transform(vm, means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    present[idx] = !(pv.z <= kNearPlane);
}

cudaError_t launch_preprocess(const PreprocessArgs& a, const Frame& f, const GeometryState& g, uint32_t* tile_count,
                              bool tbc, cudaStream_t stream) {
    const int blocks = (a.P + kPreprocessThreads - 1) / kPreprocessThreads;
    cudaError_t e = cudaMemsetAsync(g.counters, 0, sizeof(uint32_t) * 64, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(tile_count, 0, sizeof(uint32_t) * (size_t)f.grid_x * f.grid_y, stream);
    if (e != cudaSuccess) return e;
    const size_t smem = (a.colors_precomp == nullptr && a.M > 0) ? sizeof(float) * 8 * 32 * (a.M * 3 + 1) : 0;
    const bool banded = f.row0 > 0 || f.row1 < f.grid_y;
#define STP_PRE(TBC_, BANDED_)                                                                                          \
    do {                                                                                                                \
        cudaFuncSetAttribute(preprocess_kernel<TBC_, BANDED_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        preprocess_kernel<TBC_, BANDED_><<<blocks, kPreprocessThreads, smem, stream>>>(a, f, g, tile_count);            \
    } while (0)
    if (tbc) {
        if (banded) STP_PRE(true, true); else STP_PRE(true, false);
    } else {
        if (banded) STP_PRE(false, true); else STP_PRE(false, false);
    }
#undef STP_PRE
    return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, vm, present);
    return cudaGetLastError();
}

}  // namespace stp
