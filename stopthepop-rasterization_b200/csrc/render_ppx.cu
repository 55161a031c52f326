// render_ppx.cu -- per-pixel re-sorting modes: PPX_KBUFFER (forward + backward) and PPX_FULL (forward).
//
// Replaces: renderkBufferCUDA<3,W> (resorted_render.cuh:17-221), renderkBufferBackwardCUDA<3,W>
// (:223-471) and renderSortedFullCUDA<3> (:474-675).
//
// K-buffer: one thread per pixel keeps a depth-sorted window of W candidates in registers; a candidate
// enters only after passing the alpha test and the depth-along-ray >= 0 test, and the window minimum is
// blended whenever the window is full.  The tile list is staged through shared memory 256 entries at a
// time from the tile's slab (stp_slab.cuh: xy + conic/opacity + id, contiguous in list order); the depth half of the
// record (inverse covariance) is read only for candidates that survive the alpha test.
//
// Full sort: the reference sorts, for EVERY pixel, a sliding window of 4x256 list entries by the pixel's
// ray depth with a CTA-wide radix sort and lets one thread blend (256 pixels serially per CTA).  Here one
// WARP owns a pixel at a time: the 1024-entry window lives in registers (32 keys + 32 ids per lane) and is
// ordered with a register/shuffle bitonic network; the 8 warps of a CTA work on 8 pixels concurrently.
// Window semantics are the reference's (:560-672): round r adds list entries [256(r+3), 256(r+4)) (round 0
// starts from the first 1024), sorts, emits the 256 smallest in order, keeps the rest.  The sort is exact
// for lists <= 1024 entries, like the reference.
#include "stp_kernels.cuh"
#include "stp_sort.cuh"
#include "stp_slab.cuh"

namespace stp {

namespace {

constexpr float kFltMax = 3.402823466e+38f;
constexpr int kBlock = 256;

// ================================================= k-buffer ==========================================================
template <int WIN, bool BWD>
__global__ void __launch_bounds__(kBlock)
render_kbuffer_kernel(Frame f, RenderArgs a, RenderBwdArgs ab) {
    __shared__ int s_id[kBlock];
    __shared__ float2 s_xy[kBlock];
    __shared__ float4 s_co[kBlock];

    if constexpr (!BWD) {
        if (a.abort_flag != nullptr && *a.abort_flag != 0u) return;  // asynchronous forward: binning arena too small
    }
    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const uint32_t px = tile_x * 16 + (tid & 15), py = tile_y * 16 + (tid >> 4);
    const bool inside = px < (uint32_t)f.W && py < (uint32_t)f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const size_t plane = (size_t)f.W * f.H;

    const uint2* __restrict__ ranges = BWD ? ab.ranges : a.ranges;
    const float4* __restrict__ slab = BWD ? ab.slab : a.slab;
    const float2* __restrict__ means2D = BWD ? ab.means2D : a.means2D;
    const float4* __restrict__ conic_opacity = BWD ? ab.conic_opacity : a.conic_opacity;
    const float* __restrict__ colors = BWD ? ab.colors : a.colors;

    const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
    const Vec3 ray = view_ray(cam, pxf, pyf);

    const uint2 range = ranges[tile_y * f.grid_x + tile_x];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + kBlock - 1) / kBlock;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0;
    // forward: blend log slot (only the depth visualisation asks the k-buffer forward for a log)
    const uint32_t tile_lin_kb = (uint32_t)(tile_y * f.grid_x + tile_x);
    const uint32_t rec_first = tile_lin_kb * (uint32_t)a.rec_cap * 256u + (uint32_t)tid;
    uint32_t rec_idx = rec_first;
    float T_final = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, f0 = 0.f, f1 = 0.f, f2 = 0.f, bg_dot = 0.f;
    if constexpr (BWD) {
        if (inside) {
            T_final = ab.final_T[pix_id];
            g0 = ab.dL_dpix[pix_id];
            g1 = ab.dL_dpix[plane + pix_id];
            g2 = ab.dL_dpix[2 * plane + pix_id];
            f0 = ab.pixel_colors[pix_id] - T_final * f.background[0];
            f1 = ab.pixel_colors[plane + pix_id] - T_final * f.background[1];
            f2 = ab.pixel_colors[2 * plane + pix_id] - T_final * f.background[2];
        }
        bg_dot = f.background[0] * g0 + f.background[1] * g1 + f.background[2] * g2;
    }
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;

    float wd[WIN], ws[WIN];
    int wi[WIN];
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
        wd[k] = kFltMax;
        ws[k] = 0.f;
        wi[k] = -1;
    }
    int wnum = 0;

    auto blend_one = [&]() {
        if (wnum == 0) return;
        --wnum;
        const int id = wi[0];
        if constexpr (!BWD) {
            const float alpha = ws[0];
            const float test_T = fmul(T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                done = true;
                return;
            }
            C0 = ffma(fmul(__ldg(colors + 3 * id + 0), alpha), T, C0);
            C1 = ffma(fmul(__ldg(colors + 3 * id + 1), alpha), T, C1);
            C2 = ffma(fmul(__ldg(colors + 3 * id + 2), alpha), T, C2);
            T = test_T;
            if (a.blend_rec != nullptr) {
                if (rec_idx < (tile_lin_kb + 1u) * (uint32_t)a.rec_cap * 256u)
                    __stcs(a.blend_rec + rec_idx, make_uint2((uint32_t)id, __float_as_uint(alpha)));
                rec_idx += 256u;
            }
        } else {
            const float G = ws[0];
            const float4 co = __ldg(conic_opacity + id);
            const float alpha = fminf(0.99f, fmul(co.w, G));
            const float test_T = fmul(T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                done = true;
                return;
            }
            const float2 xy = __ldg(means2D + id);
            const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
            const float dchannel_dcolor = alpha * T;
            const float c0 = __ldg(colors + 3 * id + 0), c1 = __ldg(colors + 3 * id + 1), c2 = __ldg(colors + 3 * id + 2);
            C0 += c0 * alpha * T;
            C1 += c1 * alpha * T;
            C2 += c2 * alpha * T;
            const float inv_T = 1.0f / test_T;
            float dL_dalpha = (c0 - (f0 - C0) * inv_T) * g0 + (c1 - (f1 - C1) * inv_T) * g1 + (c2 - (f2 - C2) * inv_T) * g2;
            dL_dalpha *= T;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y;
            const float dG_ddely = -gdy * co.z - gdx * co.y;
            accumulate_grads(ab.grad_accum, ab.P, id, dchannel_dcolor * g0, dchannel_dcolor * g1, dchannel_dcolor * g2,
                             dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * dx * dL_dG,
                             -0.5f * gdx * dy * dL_dG, -0.5f * gdy * dy * dL_dG, G * dL_dalpha);
            T = test_T;
        }
#pragma unroll
        for (int k = 1; k < WIN; ++k) {
            wd[k - 1] = wd[k];
            ws[k - 1] = ws[k];
            wi[k - 1] = wi[k];
        }
        wd[WIN - 1] = kFltMax;
    };

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        if (__syncthreads_and(done)) break;
        const uint32_t src = range.x + r * kBlock + tid;
        if (src < range.y) {  // the alpha-test half of the tile's slab record: contiguous, no gather through the id
            float4 c0, c1;
            slab_ldg_head(slab, src, c0, c1);
            s_id[tid] = __float_as_int(c1.z);
            s_xy[tid] = make_float2(c0.x, c0.y);
            s_co[tid] = make_float4(c0.z, c0.w, c1.x, c1.y);
        }
        __syncthreads();
        const int n = min(kBlock, todo);
        for (int j = 0; !done && j < n; ++j) {
            if (wnum == WIN) blend_one();
            if (done) break;
            ++contributor;
            const int id = s_id[j];
            if (id < 0) break;  // padding entries only exist behind the last valid tile
            const float2 xy = s_xy[j];
            const float4 co = s_co[j];
            const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
            float G;
            if constexpr (!BWD) {  // forward spelling: positive factor, exp(-x)  (:165-173)
                const float pw = opacity_factor(dx, dy, co.x, co.y, co.z);
                if (pw < 0.0f) continue;
                G = expf(-pw);
            } else {  // backward spelling (:431-435)
                const float pw = gaussian_power(dx, dy, co.x, co.y, co.z);
                if (pw > 0.0f) continue;
                G = expf(pw);
            }
            const float alpha = fminf(0.99f, fmul(co.w, G));
            if (alpha < kAlphaThreshold) continue;
            float ic[6], ux, uy, uz;
            slab_ldg_inv(slab, range.x + (uint32_t)(r * kBlock + j), ic, ux, uy, uz);
            const float depth = depth_along_ray(ic, ux, uy, uz, ray);
            if (depth < 0.0f) continue;
            float ed = depth, es = BWD ? G : alpha;
            int ei = id;
#pragma unroll
            for (int k = 0; k < WIN; ++k) {
                if (ed < wd[k]) {
                    const float td = wd[k], ts = ws[k];
                    const int ti = wi[k];
                    wd[k] = ed; ws[k] = es; wi[k] = ei;
                    ed = td; es = ts; ei = ti;
                }
            }
            ++wnum;
        }
    }
    if (!done) {
        while (wnum > 0 && !done) blend_one();
    }
    if constexpr (!BWD) {
        if (inside) {
            a.final_T[pix_id] = T;
            a.n_contrib[pix_id] = contributor;
            a.out_color[pix_id] = ffma(T, f.background[0], C0);
            a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
            a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
            if (a.blend_rec != nullptr) a.blend_count[pix_id] = (rec_idx - rec_first) >> 8;
        }
    }
}

// ================================================= full sort =========================================================
// window element e (0..1023) lives in lane (e & 31), register (e >> 5): a compare-exchange with partner
// distance < 32 is a shuffle, distance >= 32 is register-local.
// Window payload v = (list index << 10) | ord, ord = the element's position in the reference's BLOCKED
// arrangement (thread*4 + item, resorted_render.cuh:560-600) -- cub::BlockRadixSort is stable in that
// order, so (key, ord) is the reference's total order and equal ray depths (they do occur) resolve
// identically.  Padding entries are (FLT_MAX, -1).  List index < 2^21 per tile.
__device__ __forceinline__ bool kv_less(float ka, int va, float kb, int vb) {
    return ka < kb || (ka == kb && (va & 1023) < (vb & 1023));
}

// full bitonic sort of 1024 (key, payload) pairs held as k[32], v[32] per lane, ascending by (key, ord).
__device__ __forceinline__ void bitonic_sort_1024(float (&k)[32], int (&v)[32], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 1024; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                switch (stride >> 5) {
#define STP_REG_STAGE(RS)                                                                       \
    case RS: {                                                                                  \
        _Pragma("unroll") for (int r = 0; r < 32; ++r) {                                        \
            if ((r & RS) == 0) {                                                                \
                const bool up = (((r << 5) | lane) & size) == 0;                                \
                const bool sw = up ? kv_less(k[r | RS], v[r | RS], k[r], v[r])                  \
                                   : kv_less(k[r], v[r], k[r | RS], v[r | RS]);                 \
                const float tk = sw ? k[r | RS] : k[r];                                         \
                const float uk = sw ? k[r] : k[r | RS];                                         \
                const int tv = sw ? v[r | RS] : v[r];                                           \
                const int uv = sw ? v[r] : v[r | RS];                                           \
                k[r] = tk; k[r | RS] = uk; v[r] = tv; v[r | RS] = uv;                           \
            }                                                                                   \
        }                                                                                       \
    } break;
                    STP_REG_STAGE(1)
                    STP_REG_STAGE(2)
                    STP_REG_STAGE(4)
                    STP_REG_STAGE(8)
                    STP_REG_STAGE(16)
#undef STP_REG_STAGE
                    default: break;
                }
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const bool up = (((r << 5) | lane) & size) == 0;
                    const float ok = __shfl_xor_sync(0xffffffffu, k[r], stride);
                    const int ov = __shfl_xor_sync(0xffffffffu, v[r], stride);
                    // the lower lane keeps the smaller element when sorting upwards
                    const bool take = (lower == up) ? kv_less(ok, ov, k[r], v[r]) : kv_less(k[r], v[r], ok, ov);
                    k[r] = take ? ok : k[r];
                    v[r] = take ? ov : v[r];
                }
            }
        }
    }
}

// ---- fast path for tiles of at most 1024 instances ------------------------------------------------------------------
// For such a tile the reference's sliding window holds the whole list from the first round on, i.e. the blending
// order is the exact sort of ALL entries by the pixel's ray depth -- and entries that fail the alpha test never
// blend.  So it is enough to sort the SURVIVORS of the alpha test (typically a tenth of the list): same image,
// same final_T, and n_contrib (= rank of the last contributor among all entries + 1) follows from one counting pass
// over the keys.  Exact key ties, whose order in the reference depends on the history of its window, are not
// resolved here: a pixel with a tie among its survivors, a tie of its last contributor with any entry, or more
// than kFastSurv survivors is marked (n_contrib = ~0) and left to render_full_kernel below.
//
// Data movement: the tile's slab (64-byte records in list order, written by the tile-sort epilogue, stp_slab.cuh) is
// pulled into shared memory ONCE per tile with a single bulk-async copy (TMA, completion on an mbarrier) and then
// serves all 256 pixels: every (pixel, entry) evaluation is four conflict-free 128-bit shared-memory loads instead of
// 76 bytes of global gathers through the instance's Gaussian id (the reference re-fetches per pixel,
// resorted_render.cuh:598-600).  One 512-thread CTA per SM: 16 warps, each owning one pixel row of the tile.
constexpr int kFastSurv = 256;
constexpr int kFastWarps = 16;
constexpr int kFastThreads = kFastWarps * 32;
constexpr uint32_t kSlowMark = 0xFFFFFFFFu;
struct FullFastShared {
    float4 slab[1024 * kSlabChunks];                  // 64 KB, destination of the bulk copy
    uint32_t key[kFastWarps][1024];                   // order-preserving integer image of every entry's ray depth, per warp
    unsigned long long surv[kFastWarps][kFastSurv];   // (key << 32 | list position) of the alpha-test survivors
    uint64_t bar;
};

__device__ __forceinline__ uint32_t sortable_bits(float x) {  // monotone float -> uint32 (-0 folded into +0)
    const uint32_t b = __float_as_uint(x + 0.0f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// alpha of slab record idx at the pixel, the forward spelling of resorted_render.cuh:165-173; false = rejected
__device__ __forceinline__ bool slab_alpha(const float4* __restrict__ slab, int idx, uint32_t first, float pxf, float pyf,
                                           float& alpha, int& id) {
    float4 c0, c1;
    slab_load_head(slab + 4 * idx, first + (uint32_t)idx, c0, c1);
    id = __float_as_int(c1.z);
    const float dx = fsub(c0.x, pxf), dy = fsub(c0.y, pyf);
    const float pw = opacity_factor(dx, dy, c0.z, c0.w, c1.x);
    if (pw < 0.0f) return false;
    alpha = fminf(0.99f, fmul(c1.y, expf(-pw)));
    return !(alpha < kAlphaThreshold);
}

// sorts the S survivors of one pixel (E*32 >= S) and blends them front to back; returns false if a key tie was found.
// The transmittance chain is sequential (it decides where the pixel stops, and with it n_contrib, final_T and the blend
// log, all bit-exact); the colour sum of a group of 32 blends is a warp tree reduction.
template <int E>
__device__ __forceinline__ bool full_fast_sort_blend(const FullFastShared& sh, int warp, int lane, int S, uint32_t first,
                                                     float pxf, float pyf, const RenderArgs& a, bool logging,
                                                     uint32_t rec_first, float& T, float& C0, float& C1, float& C2,
                                                     uint32_t& last_key, bool& have_last, uint32_t& nrec) {
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int e = r * 32 + lane;
        v[r] = e < S ? sh.surv[warp][e] : ~0ull;
    }
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) reg_stage<E>(v, j, k, 0, lane);
    }
    // key ties among neighbours of the sorted sequence?
    bool tie = false;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const uint32_t mine = (uint32_t)(v[r] >> 32);
        uint32_t next = __shfl_down_sync(0xffffffffu, mine, 1);
        const uint32_t wrap = __shfl_sync(0xffffffffu, (uint32_t)(v[(r + 1 < E) ? r + 1 : r] >> 32), 0);
        if (lane == 31) next = (r + 1 < E) ? wrap : 0xFFFFFFFFu;
        tie |= (r * 32 + lane + 1 < S) && mine == next;
    }
    if (__any_sync(0xffffffffu, tie)) return false;
    bool done = false;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        if (done || r * 32 >= S) break;
        const int e = r * 32 + lane;
        const uint32_t mykey = (uint32_t)(v[r] >> 32);
        float alpha = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;
        const int idx = (int)((uint32_t)v[r] & 1023u);  // tile-local list position: what the blend log stores
        if (e < S) {
            int id;
            slab_alpha(sh.slab, idx, first, pxf, pyf, alpha, id);  // passed before: same bits
            const float4 cr = __ldg(a.slab_rgb + first + (uint32_t)idx);
            w0 = fmul(cr.x, alpha);
            w1 = fmul(cr.y, alpha);
            w2 = fmul(cr.z, alpha);
        }
        const int cnt = min(32, S - r * 32);
        float myT = 0.f;
        int stop = cnt;
        for (int l = 0; l < cnt; ++l) {
            const float al = __shfl_sync(0xffffffffu, alpha, l);
            const float test_T = fmul(T, fsub(1.0f, al));
            if (test_T < kTThreshold) {
                done = true;
                stop = l;
                break;
            }
            myT = (lane == l) ? T : myT;
            T = test_T;
        }
        if (stop > 0) {
            float t0 = w0 * myT, t1 = w1 * myT, t2 = w2 * myT;  // lanes >= stop hold myT = 0
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t0 += __shfl_xor_sync(0xffffffffu, t0, o);
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            }
            C0 += t0;
            C1 += t1;
            C2 += t2;
            last_key = __shfl_sync(0xffffffffu, mykey, stop - 1);
            have_last = true;
            if (logging) {
                if (lane < stop && nrec + (uint32_t)lane < (uint32_t)a.rec_cap)
                    __stcs(a.blend_rec + rec_first + (nrec + (uint32_t)lane) * 256u, make_uint2((uint32_t)idx, __float_as_uint(alpha)));
                nrec += (uint32_t)stop;
            }
        }
    }
    return true;
}

__global__ void __launch_bounds__(kFastThreads, 1)
render_full_fast_kernel(Frame f, RenderArgs a) {
    extern __shared__ __align__(128) unsigned char smem_full[];
    FullFastShared& sh = *reinterpret_cast<FullFastShared*>(smem_full);
    if (a.abort_flag != nullptr && *a.abort_flag != 0u) return;  // asynchronous forward: binning arena too small
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const size_t plane = (size_t)f.W * f.H;
    const uint32_t tile_lin = (uint32_t)(tile_y * f.grid_x + tile_x);
    const uint2 range = a.ranges[tile_lin];
    const int n = (int)(range.y - range.x);
    if (n > 1024) return;  // render_full_kernel emulates the sliding window for long lists
    // the tile's slab: one bulk-async copy, everybody waits on its mbarrier
    if (tid == 0) {
        mbar_init(&sh.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && n > 0) {
        mbar_expect_tx(&sh.bar, (uint32_t)n * kSlabRecordBytes);
        bulk_g2s(sh.slab, a.slab + (size_t)kSlabChunks * range.x, (uint32_t)n * kSlabRecordBytes, &sh.bar);
    }
    const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
    const bool logging = a.blend_rec != nullptr;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t py = tile_y * 16 + warp;
    if (n > 0) mbar_wait(&sh.bar, 0);
    if (!(py < (uint32_t)f.H)) return;  // warp-uniform; nothing is pending any more

    for (int pi = 0; pi < 16; ++pi) {
        const uint32_t px = tile_x * 16 + pi;
        if (!(px < (uint32_t)f.W)) break;  // warp-uniform
        const uint32_t pix_id = (uint32_t)f.W * py + px;
        const float pxf = (float)px, pyf = (float)py;
        const Vec3 ray = view_ray_xloop(cam, pxf, pyf);
        __syncwarp();  // the previous pixel's reads of this warp's shared-memory rows are complete
        int S = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane;
            bool accept = false;
            uint32_t skey = 0xFFFFFFFFu;
            if (idx < n) {
                const uint32_t j = range.x + (uint32_t)idx;
                float4 c0, c1, c2, c3;
                slab_load_head(sh.slab + 4 * idx, j, c0, c1);
                slab_load_tail(sh.slab + 4 * idx, j, c2, c3);
                const float ic[6] = {c1.w, c2.x, c2.y, c2.z, c2.w, c3.x};
                skey = sortable_bits(depth_along_ray(ic, c3.y, c3.z, c3.w, ray));
                sh.key[warp][idx] = skey;
                const float dx = fsub(c0.x, pxf), dy = fsub(c0.y, pyf);
                const float pw = opacity_factor(dx, dy, c0.z, c0.w, c1.x);
                if (!(pw < 0.0f)) accept = !(fminf(0.99f, fmul(c1.y, expf(-pw))) < kAlphaThreshold);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, accept);
            const int slot = S + __popc(m & lt_mask);
            if (accept && slot < kFastSurv)
                sh.surv[warp][slot] = ((unsigned long long)skey << 32) | (unsigned long long)idx;
            S += __popc(m);
        }
        __syncwarp();
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
        uint32_t last_key = 0, nrec = 0;
        bool have_last = false;
        const uint32_t rec_first = tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)(warp * 16 + pi);
        bool ok = S <= kFastSurv;
        if (ok) {
#define STP_FAST(E_) full_fast_sort_blend<E_>(sh, warp, lane, S, range.x, pxf, pyf, a, logging, rec_first, T, C0, C1, C2, last_key, have_last, nrec)
            if (S <= 32) ok = STP_FAST(1);
            else if (S <= 64) ok = STP_FAST(2);
            else if (S <= 128) ok = STP_FAST(4);
            else ok = STP_FAST(8);
#undef STP_FAST
        }
        uint32_t last_contributor = 0;
        if (ok && have_last) {
            // rank of the last contributor among ALL entries of the tile
            int less = 0, equal = 0;
            for (int idx = lane; idx < n; idx += 32) {
                const uint32_t kk = sh.key[warp][idx];
                less += kk < last_key;
                equal += kk == last_key;
            }
            less = __reduce_add_sync(0xffffffffu, less);
            equal = __reduce_add_sync(0xffffffffu, equal);
            if (equal != 1) ok = false;
            last_contributor = (uint32_t)less + 1u;
        }
        if (lane == 0) {
            if (!ok) {
                a.n_contrib[pix_id] = kSlowMark;
            } else {
                a.final_T[pix_id] = T;
                a.n_contrib[pix_id] = last_contributor;
                a.out_color[pix_id] = ffma(T, f.background[0], C0);
                a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
                a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
                if (logging) {
                    a.blend_count[pix_id] = nrec;
                    if (nrec > (uint32_t)a.rec_cap) atomicAdd(a.log_overflow, 1u);
                }
            }
        }
    }
}

// ---- exact emulation of the reference's sliding window: long lists, and the pixels the fast path gave up on -------------
__global__ void __launch_bounds__(kBlock)
render_full_kernel(Frame f, RenderArgs a) {
    if (a.abort_flag != nullptr && *a.abort_flag != 0u) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const size_t plane = (size_t)f.W * f.H;
    const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
    const uint2 range = a.ranges[tile_y * f.grid_x + tile_x];
    const int n = (int)(range.y - range.x);
    const int rounds = (n + 255) / 256;

    const bool logging = a.blend_rec != nullptr;
    const uint32_t tile_lin = (uint32_t)(tile_y * f.grid_x + tile_x);
    // warp w renders the 32 pixels of rows 2w, 2w+1 of the tile, one after the other
    for (int pi = 0; pi < 32; ++pi) {
        const uint32_t px = tile_x * 16 + (pi & 15), py = tile_y * 16 + warp * 2 + (pi >> 4);
        if (!(px < (uint32_t)f.W && py < (uint32_t)f.H)) continue;  // warp-uniform
        const uint32_t pix_id = (uint32_t)f.W * py + px;
        if (n <= 1024 && a.n_contrib[pix_id] != kSlowMark) continue;  // done by render_full_fast_kernel (warp-uniform)
        const float pxf = (float)px, pyf = (float)py;
        const Vec3 ray = view_ray_xloop(cam, pxf, pyf);

        float k[32];
        int v[32];
        // the first 768 list entries; slot r*32+lane <- list entry r*32+lane (any placement is equivalent
        // up to exact key ties), the last 256 slots are refilled every round
        auto load_slot = [&](int r, int idx, int ord) {
            float key = kFltMax;
            int payload = -1;
            if (idx < n) {
                float ic[6], ux, uy, uz;
                slab_ldg_inv(a.slab, range.x + (uint32_t)idx, ic, ux, uy, uz);
                key = depth_along_ray(ic, ux, uy, uz, ray);
                payload = (idx << 10) | ord;
            }
            k[r] = key;
            v[r] = payload;
        };
        // list entry idx = i*256 + t (i < 3) sits at blocked position 4t + i + 1 in the reference (:560-575)
#pragma unroll
        for (int r = 0; r < 24; ++r) {
            const int idx = r * 32 + lane;
            load_slot(r, idx, 4 * (idx & 255) + (idx >> 8) + 1);
        }

        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
        uint32_t last_contributor = 0;
        bool done = false;
        int todo = n;
        // blend log of this pixel (row-major position warp*32+pi inside the tile): blend_rec[tile][k][position]
        const uint32_t rec_first = tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)(warp * 32 + pi);
        uint32_t nrec = 0;
        for (int rd = 0; rd < rounds && !done; ++rd, todo -= 256) {
#pragma unroll
            for (int r = 24; r < 32; ++r) {  // the round's 256 new entries take item 0 of every thread (:588-600)
                const int j = (r - 24) * 32 + lane;
                load_slot(r, (rd + 3) * 256 + j, 4 * j);
            }
            bitonic_sort_1024(k, v, lane);
            // the 256 smallest are elements 0..255 = registers 0..7.  Every lane evaluates alpha of ITS element;
            // the accepted ones are then blended in order (sequential transmittance, like the reference's thread 0)
            const int lim = min(256, todo);
            bool stop = false;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (!stop && !done) {
                    const int e = r * 32 + lane;
                    float alpha = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
                    bool accept = false;
                    int my_id = -1;
                    if (e < lim && v[r] >= 0) {
                        float4 h0, h1;
                        slab_ldg_head(a.slab, range.x + (uint32_t)(v[r] >> 10), h0, h1);
                        my_id = v[r] >> 10;  // tile-local list position: what the blend log stores
                        const float dx = fsub(h0.x, pxf), dy = fsub(h0.y, pyf);
                        const float pw = opacity_factor(dx, dy, h0.z, h0.w, h1.x);
                        if (!(pw < 0.0f)) {
                            alpha = fminf(0.99f, fmul(h1.y, expf(-pw)));
                            accept = !(alpha < kAlphaThreshold);
                        }
                        if (accept) {
                            const float4 cr = __ldg(a.slab_rgb + range.x + (uint32_t)my_id);
                            c0 = cr.x;
                            c1 = cr.y;
                            c2 = cr.z;
                        }
                    }
                    uint32_t m = __ballot_sync(0xffffffffu, accept);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const float al = __shfl_sync(0xffffffffu, alpha, l);
                        const float b0 = __shfl_sync(0xffffffffu, c0, l), b1 = __shfl_sync(0xffffffffu, c1, l),
                                    b2 = __shfl_sync(0xffffffffu, c2, l);
                        const float test_T = fmul(T, fsub(1.0f, al));
                        if (test_T < kTThreshold) {
                            done = true;
                            break;
                        }
                        C0 = ffma(fmul(b0, al), T, C0);
                        C1 = ffma(fmul(b1, al), T, C1);
                        C2 = ffma(fmul(b2, al), T, C2);
                        T = test_T;
                        last_contributor = (uint32_t)(rd * 256 + r * 32 + l + 1);
                        if (logging) {
                            if (lane == l && nrec < (uint32_t)a.rec_cap)
                                __stcs(a.blend_rec + rec_first + nrec * 256u, make_uint2((uint32_t)my_id, __float_as_uint(al)));
                            ++nrec;
                        }
                    }
                    if ((r + 1) * 32 >= lim) stop = true;
                }
            }
            // keep elements 256..1023 as slots 0..767 of the next round
#pragma unroll
            for (int r = 0; r < 24; ++r) {  // rank rk = 256 + e' goes to thread rk % 256, item rk / 256 (striped, :664-672)
                const int e2 = r * 32 + lane;
                k[r] = k[r + 8];
                v[r] = v[r + 8] < 0 ? -1 : ((v[r + 8] & ~1023) | (4 * (e2 & 255) + (e2 >> 8) + 1));
            }
        }
        if (lane == 0) {
            a.final_T[pix_id] = T;
            a.n_contrib[pix_id] = last_contributor;
            a.out_color[pix_id] = ffma(T, f.background[0], C0);
            a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
            a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
            if (logging) {
                a.blend_count[pix_id] = nrec;
                if (nrec > (uint32_t)a.rec_cap) atomicAdd(a.log_overflow, 1u);
            }
        }
    }
}

template <bool BWD>
cudaError_t dispatch_kbuffer(const Frame& f, int window, const RenderArgs& a, const RenderBwdArgs& ab, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
#define STP_KB(W_) render_kbuffer_kernel<W_, BWD><<<grid, kBlock, 0, stream>>>(f, a, ab)
    // window rounding of forward.cu:410-425 / backward.cu:714-731
    if (window <= 1) STP_KB(1);
    else if (window <= 2) STP_KB(2);
    else if (window <= 4) STP_KB(4);
    else if (window <= 8) STP_KB(8);
    else if (window <= 12) STP_KB(12);
    else if (window <= 16) STP_KB(16);
    else if (window <= 20) STP_KB(20);
    else STP_KB(24);
#undef STP_KB
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_render_kbuffer_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream) {
    RenderBwdArgs dummy{};
    return dispatch_kbuffer<false>(f, s.q_head, a, dummy, stream);
}
cudaError_t launch_render_kbuffer_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream) {
    RenderArgs dummy{};
    return dispatch_kbuffer<true>(f, s.q_head, dummy, a, stream);
}
cudaError_t launch_render_full_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    // per launch: the attribute belongs to the current device, a process may drive several GPUs
    cudaFuncSetAttribute(render_full_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FullFastShared));
    render_full_fast_kernel<<<grid, kFastThreads, sizeof(FullFastShared), stream>>>(f, a);
    render_full_kernel<<<grid, kBlock, 0, stream>>>(f, a);
    return cudaGetLastError();
}

// PPX_FULL backward: the reference has none (backward.cu:733-736).  With the blend log of the forward pass it is the
// same replay as for the other modes: the per-pixel sort order is held fixed, like in the k-buffer / hierarchical backward.
cudaError_t launch_render_full_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream) {
    return launch_blend_replay_bwd(f, a, 2, false, stream);
}

}  // namespace stp
