// render_global.cu -- GLOBAL sort mode: front-to-back alpha blending of each tile's sorted list
// and the matching back-to-front gradient sweep.
//
// Replaces: renderCUDA<3,false> forward (forward.cu:234-366) and renderCUDA<3> backward
// (backward.cu:437-595).
//
// One CTA per 16x16 tile, one thread per pixel; warp w owns the 8x4-pixel strip (w&1, w>>1) of the tile.
// The tile's list is streamed through shared memory in slabs of 256 entries {xy, conic+opacity, rgb}.
//
// Both render kernels are FP32-issue bound, not HBM bound (ncu, profiles/r01a_*: issue slots 90 % busy,
// DRAM 1 %): the reference evaluates every (pixel, list entry) pair, 256 evaluations per entry and tile,
// although a typical splat only reaches a few dozen pixels of the tile.  Here the thread that stages
// entry j of a slab also computes, once, WHICH of the 8 strips the entry can reach: the axis-aligned
// bounding box of the ellipse { q(d) <= ln(255*opacity) } (inflated, so the test is conservative with
// respect to the float arithmetic of the per-pixel evaluation) against the strip rectangles -> an
// 8-bit mask per entry.  A warp then only visits the entries whose mask names its strip (ballot over
// 32 masks + find-first-set walk, order preserved).  A skipped evaluation could only have ended in the
// "alpha < 1/255 -> continue" branch, which has no side effect besides the list-position counter (kept
// arithmetically), so final_T, n_contrib and the image are bit-identical to evaluating everything.
// All per-(pixel,Gaussian) arithmetic is the rounding-pinned sequence of stp_math.cuh.
//
// Backward: the nine per-Gaussian gradient terms are reduced across the warp with a recursive-halving
// exchange (16 shuffles instead of 45) and accumulated per CTA in shared memory; one global float
// atomic per (tile, Gaussian, term) remains instead of one per (pixel, Gaussian, term) in the reference
// (backward.cu:561,583-592).
#include "stp_kernels.cuh"

namespace stp {

namespace {

constexpr int kTile = 16;
constexpr int kBlock = kTile * kTile;

// Which of the 8 strips (bit = strip index = warp index; strip (sx,sy) = pixels [8sx,8sx+7] x [4sy,4sy+3]
// of the tile) can contain a pixel with alpha >= 1/255 for this Gaussian?  Conservative: real q(d) <= thr
// implies |dx| <= sqrt(2 thr C/det), |dy| <= sqrt(2 thr A/det); thr is inflated by 0.1 % + 0.01 and the
// half-widths by 0.01 px, orders of magnitude more than the rounding error of the evaluated power.
__device__ __forceinline__ uint32_t strip_mask(float2 xy, float4 co, float tile_x0, float tile_y0) {
    const float det = co.x * co.z - co.y * co.y;
    if (!(det > 0.0f) || !(co.x > 0.0f) || !(co.z > 0.0f) || !(co.w > 0.0f)) return 0xffu;
    const float thr = __logf(255.0f * co.w) * 1.001f + 0.01f;
    if (!(thr > 0.0f)) return 0xffu;
    const float s = 2.0f * thr / det;
    const float hx = sqrtf(s * co.z) * 1.0001f + 0.01f, hy = sqrtf(s * co.x) * 1.0001f + 0.01f;
    if (!(hx < 1e9f) || !(hy < 1e9f)) return 0xffu;
    const float lx = xy.x - hx - tile_x0, ux = xy.x + hx - tile_x0;
    const float ly = xy.y - hy - tile_y0, uy = xy.y + hy - tile_y0;
    uint32_t col = 0, mask = 0;
    col |= (ux >= 0.0f && lx <= 7.0f) ? 1u : 0u;
    col |= (ux >= 8.0f && lx <= 15.0f) ? 2u : 0u;
#pragma unroll
    for (int sy = 0; sy < 4; ++sy)
        if (uy >= 4.0f * sy && ly <= 4.0f * sy + 3.0f) mask |= col << (2 * sy);
    return mask;
}

// minBlocks = 1 lets ptxas keep the slab loop's addresses in registers: 0.35 -> 0.32 ms at C2 (A/B on B200; 6 and 8: 0.35)
#ifndef STP_GLOBAL_FWD_MINB
#define STP_GLOBAL_FWD_MINB 1
#endif
template <bool LOG>  // LOG: write the blend log (off by default for GLOBAL; the plain instantiation carries none of its code)
__global__ void __launch_bounds__(kBlock, STP_GLOBAL_FWD_MINB)
render_global_fwd_kernel(Frame f, RenderArgs a) {
    __shared__ float2 s_xy[kBlock];
    __shared__ float4 s_co[kBlock];
    __shared__ float s_rgb[3][kBlock];
    __shared__ uint32_t s_mask[kBlock];
    __shared__ uint32_t s_id[kBlock];  // only used when the blend log is written

    if (a.abort_flag != nullptr && *a.abort_flag != 0u) return;  // asynchronous forward: binning arena too small (api.cu)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const uint32_t px = tile_x * kTile + (warp & 1) * 8 + (lane & 7), py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)f.W && py < (uint32_t)f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

    const uint32_t tile_lin = (uint32_t)(tile_y * f.grid_x + tile_x);
    const uint2 range = a.ranges[tile_lin];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + kBlock - 1) / kBlock;
    // blend log (training steps): slot of this pixel's next blend, blend_rec[tile][k][thread]; 32-bit index (checked by the host)
    constexpr bool logging = LOG;
    const uint32_t rec_first = tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)tid;
    const uint32_t rec_end = (tile_lin + 1u) * (uint32_t)a.rec_cap * 256u;
    uint32_t rec_idx = rec_first;

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last_contributor = 0;

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        if (__syncthreads_count(done) == kBlock) break;
        const uint32_t src = range.x + r * kBlock + tid;
        uint32_t mask = 0;
        if (src < range.y) {
            const uint32_t id = a.point_list[src];
            const float2 xy = a.means2D[id];
            const float4 co = a.conic_opacity[id];
            s_xy[tid] = xy;
            s_co[tid] = co;
            s_rgb[0][tid] = a.colors[3 * id + 0];
            s_rgb[1][tid] = a.colors[3 * id + 1];
            s_rgb[2][tid] = a.colors[3 * id + 2];
            if constexpr (logging) s_id[tid] = id;
            mask = strip_mask(xy, co, tile_x0, tile_y0);
        }
        s_mask[tid] = mask;
        __syncthreads();
        const int n = min(kBlock, todo);
        for (int c = 0; c * 32 < n; ++c) {
            if (__all_sync(0xffffffffu, done)) break;
            uint32_t m = __ballot_sync(0xffffffffu, (s_mask[c * 32 + lane] >> warp) & 1u);
            while (m) {
                const int j = c * 32 + __ffs(m) - 1;
                m &= m - 1;
                if (done) continue;
                const float2 xy = s_xy[j];
                const float4 co = s_co[j];
                const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
                const float power = gaussian_power(dx, dy, co.x, co.y, co.z);
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, fmul(co.w, expf(power)));
                if (alpha < kAlphaThreshold) continue;
                const float test_T = fmul(T, fsub(1.0f, alpha));
                if (test_T < kTThreshold) {
                    done = true;
                    continue;
                }
                C0 = ffma(T, fmul(alpha, s_rgb[0][j]), C0);
                C1 = ffma(T, fmul(alpha, s_rgb[1][j]), C1);
                C2 = ffma(T, fmul(alpha, s_rgb[2][j]), C2);
                T = test_T;
                last_contributor = (uint32_t)(r * kBlock + j + 1);
                if constexpr (logging) {
                    if (rec_idx < rec_end) __stcs(a.blend_rec + rec_idx, make_uint2(s_id[j], __float_as_uint(alpha)));
                    rec_idx += 256u;
                }
            }
        }
    }

    if (inside) {
        a.final_T[pix_id] = T;
        a.n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)f.W * f.H;
        a.out_color[pix_id] = ffma(T, f.background[0], C0);
        a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
        a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
        if constexpr (logging) {
            const uint32_t nrec = (rec_idx - rec_first) >> 8;
            a.blend_count[pix_id] = nrec;
            if (nrec > (uint32_t)a.rec_cap) atomicOr(a.tile_flags + tile_lin, 1u);
        }
    }
}

constexpr int kRedBatch = 2;  // entries whose terms are reduced together (2 x 9 rows; 20.7 KB of panels per CTA)

__global__ void __launch_bounds__(kBlock)  // A/B on B200: explicit minBlocks 1 / 5 / 6 are all slower (0.94 / 0.91 / 0.99 vs 0.89 ms)
render_global_bwd_kernel(Frame f, RenderBwdArgs a) {
    __shared__ uint32_t s_id[kBlock];
    __shared__ float2 s_xy[kBlock];
    __shared__ float4 s_co[kBlock];
    __shared__ float s_rgb[3][kBlock];
    __shared__ uint32_t s_mask[kBlock];
    __shared__ float s_acc[kBlock][9];  // per-slab gradient accumulators (stride 9: the nine terms of one entry hit nine banks)
    // Cross-lane reduction of the nine gradient terms, transposed through shared memory: every lane parks its nine
    // values of an entry in a [term][lane] panel (row stride 36 floats: conflict-free 128-bit reads); once kRedBatch
    // entries are parked, 9 * kRedBatch lanes each sum one row of 32 and add it to the entry's accumulator.  9 stores +
    // ~13 instructions per entry instead of the ~60 of a register-shuffle reduction of nine terms.
    __shared__ __align__(16) float s_red[kBlock / 32][kRedBatch][9][36];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const uint32_t px = tile_x * kTile + (warp & 1) * 8 + (lane & 7), py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)f.W && py < (uint32_t)f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const float tile_x0 = (float)(tile_x * kTile), tile_y0 = (float)(tile_y * kTile);

    // with a blend log the replay kernel has done every tile whose pixels all fit the log; this list-driven sweep
    // only handles the tiles that overflowed
    if (a.blend_rec != nullptr && a.tile_flags[tile_y * f.grid_x + tile_x] == 0u) return;

    const uint2 range = a.ranges[tile_y * f.grid_x + tile_x];
    int todo = (int)(range.y - range.x);
    const int total = todo;
    const int rounds = (todo + kBlock - 1) / kBlock;

    const float T_final = inside ? a.final_T[pix_id] : 0.f;
    float T = T_final;
    const uint32_t last_contributor = inside ? a.n_contrib[pix_id] : 0u;
    // nothing behind the last contributor of any pixel of the tile can receive a gradient
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last_contributor);

    const size_t plane = (size_t)f.W * f.H;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (inside) {
        g0 = a.dL_dpix[pix_id];
        g1 = a.dL_dpix[plane + pix_id];
        g2 = a.dL_dpix[2 * plane + pix_id];
    }
    const float bg_dot = f.background[0] * g0 + f.background[1] * g1 + f.background[2] * g2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;
    // reduction panels of this warp: rows parked so far and the slab slots they belong to
    float (*const red)[9][36] = s_red[warp];
    int parked = 0, parked_j0 = 0, parked_j1 = 0;
    auto flush_parked = [&]() {
        __syncwarp();
        if (lane < 9 * parked) {
            const int b = lane >= 9 ? 1 : 0, term = lane - 9 * b;
            const float4* const row = reinterpret_cast<const float4*>(red[b][term]);
            float4 s4 = row[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                const float4 t = row[q];
                s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
            }
            const float tot = (s4.x + s4.y) + (s4.z + s4.w);
            if (tot != 0.0f) atomicAdd(&s_acc[b ? parked_j1 : parked_j0][term], tot);
        }
        parked = 0;
        __syncwarp();
    };

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        __syncthreads();
        // slab r holds list positions total-1-(r*256+t), t = 0..255 (back to front)
        const int progress = r * kBlock + tid;
        uint32_t mask = 0;
        if (range.x + progress < range.y) {
            const uint32_t id = a.point_list[range.y - progress - 1];
            const float2 xy = a.means2D[id];
            const float4 co = a.conic_opacity[id];
            s_id[tid] = id;
            s_xy[tid] = xy;
            s_co[tid] = co;
            s_rgb[0][tid] = a.colors[3 * id + 0];
            s_rgb[1][tid] = a.colors[3 * id + 1];
            s_rgb[2][tid] = a.colors[3 * id + 2];
            mask = strip_mask(xy, co, tile_x0, tile_y0);
        }
        s_mask[tid] = mask;
#pragma unroll
        for (int k = 0; k < 9; ++k) s_acc[tid][k] = 0.f;
        __syncthreads();
        const int n = min(kBlock, todo);
        for (int c = 0; c * 32 < n; ++c) {
            // list position (0-based) of entry j of this slab: total-1-(r*256+j); it contributes to a pixel iff
            // position < last_contributor of that pixel
            const int pos_first = total - 1 - (r * kBlock + c * 32);
            if (pos_first - 31 >= (int)warp_last) continue;  // the whole chunk lies behind every pixel's last contributor
            uint32_t m = __ballot_sync(0xffffffffu, (s_mask[c * 32 + lane] >> warp) & 1u);
            while (m) {
                const int j = c * 32 + __ffs(m) - 1;
                m &= m - 1;
                const uint32_t contributor = (uint32_t)(total - 1 - (r * kBlock + j));
                bool hit = inside && contributor < last_contributor;
                float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
                float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
                if (hit) {
                    const float2 xy = s_xy[j];
                    co = s_co[j];
                    dx = fsub(xy.x, pxf);
                    dy = fsub(xy.y, pyf);
                    const float power = gaussian_power(dx, dy, co.x, co.y, co.z);
                    hit = !(power > 0.0f);
                    if (hit) {
                        G = expf(power);
                        alpha = fminf(0.99f, fmul(co.w, G));
                        hit = !(alpha < kAlphaThreshold);
                    }
                }
                if (!__any_sync(0xffffffffu, hit)) continue;
                float v[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] = 0.f;
                if (hit) {
                    // one reciprocal (MUFU.RCP, 1 ulp) shared by the two divisions of backward.cu:541,571
                    const float inv_1ma = __fdividef(1.f, 1.f - alpha);
                    T = T * inv_1ma;
                    const float dchannel_dcolor = alpha * T;
                    const float c0 = s_rgb[0][j], c1 = s_rgb[1][j], c2 = s_rgb[2][j];
                    acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                    acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                    acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                    lc0 = c0;
                    lc1 = c1;
                    lc2 = c2;
                    float dL_dalpha = (c0 - acc0) * g0 + (c1 - acc1) * g1 + (c2 - acc2) * g2;
                    v[0] = dchannel_dcolor * g0;
                    v[1] = dchannel_dcolor * g1;
                    v[2] = dchannel_dcolor * g2;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final * inv_1ma) * bg_dot;
                    const float dL_dG = co.w * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co.x - gdy * co.y;
                    const float dG_ddely = -gdy * co.z - gdx * co.y;
                    v[3] = dL_dG * dG_ddelx * ddelx_dx;
                    v[4] = dL_dG * dG_ddely * ddely_dy;
                    v[5] = -0.5f * gdx * dx * dL_dG;
                    v[6] = -0.5f * gdx * dy * dL_dG;
                    v[7] = -0.5f * gdy * dy * dL_dG;
                    v[8] = G * dL_dalpha;
                }
#pragma unroll
                for (int k = 0; k < 9; ++k) red[parked][k][lane] = v[k];
                if (parked == 0) parked_j0 = j; else parked_j1 = j;
                if (++parked == kRedBatch) flush_parked();
            }
        }
        if (parked) flush_parked();
        __syncthreads();
        if (tid < n) {
            const uint32_t id = s_id[tid];
            const float c0 = s_acc[tid][0], c1 = s_acc[tid][1], c2 = s_acc[tid][2];
            const float m0 = s_acc[tid][3], m1 = s_acc[tid][4];
            const float k0 = s_acc[tid][5], k1 = s_acc[tid][6], k2 = s_acc[tid][7];
            const float o = s_acc[tid][8];
            float* const acc = a.grad_accum;
            if (k0 != 0.f || k1 != 0.f || k2 != 0.f || o != 0.f) red_add_v4(acc + 4 * (size_t)id, k0, k1, k2, o);
            if (m0 != 0.f || m1 != 0.f || c0 != 0.f || c1 != 0.f) red_add_v4(acc + 4 * ((size_t)a.P + id), m0, m1, c0, c1);
            if (c2 != 0.f) atomicAdd(acc + 8 * (size_t)a.P + id, c2);
        }
    }
}

}  // namespace

cudaError_t launch_render_global_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    if (a.blend_rec != nullptr) {
        cudaError_t e = cudaMemsetAsync(a.tile_flags, 0, sizeof(uint32_t) * (size_t)f.grid_x * f.grid_y, stream);
        if (e != cudaSuccess) return e;
    }
    if (a.blend_rec != nullptr) render_global_fwd_kernel<true><<<grid, kBlock, 0, stream>>>(f, a);
    else render_global_fwd_kernel<false><<<grid, kBlock, 0, stream>>>(f, a);
    return cudaGetLastError();
}

cudaError_t launch_render_global_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    if (a.blend_rec != nullptr) {
        cudaError_t e = launch_blend_replay_bwd(f, a, 0, true, stream);
        if (e != cudaSuccess) return e;
    }
    render_global_bwd_kernel<<<grid, kBlock, 0, stream>>>(f, a);
    return cudaGetLastError();
}

}  // namespace stp
