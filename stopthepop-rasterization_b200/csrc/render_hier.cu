// render_hier.cu -- HIER sort mode: three-level (4x4 tail / 2x2 mid / per-pixel head) streaming depth
// re-sort fused with front-to-back alpha blending, forward and backward.
//
// Replaces: sortGaussiansRayHierarchicaEvaluation (hierarchical_render.cuh:207-935) with its forward
// (:939-1035) and backward (:1038-1175) kernels.
//
// What is kept from the reference is the QUEUE SEMANTICS (which Gaussian is popped when a finite
// queue is full decides the image):
//   * the tile list is consumed 32 entries at a time; per 4x4 block the 32 new entries get a depth on
//     the block-centre ray (optionally after the 4x4 contribution cull), are sorted and merged into
//     the tail queue (resident entries first on ties); while the tail holds more than 32 entries its
//     16 smallest leave in groups of 4;
//   * every 2x2 quad re-evaluates each group on its own ray, rank-sorts it (ties by position) and merges
//     it into its mid queue (resident first on ties); when the queue holds more than MID-4 entries the
//     4 smallest go to the four pixels of the quad;
//   * every pixel first blends its head minimum if the head is full, then evaluates the entry on its
//     own ray (depth < 0, power > 0, alpha < 1/255 reject) and inserts it by strict '<';
//   * drain order tail -> mid -> head.
// How it is done is new: one warp owns two 4x4 blocks (one per half-warp), all queue manipulation is
// rank based (every lane computes the final position of "its" entries with branch-free counting /
// binary search and scatters them once) instead of compare-exchange networks with a barrier per
// stage; queues are addressed through base offsets so popping never moves data; warps of one tile
// run independently (no CTA barrier in the streaming loop), so a finished 8x4 region retires early.
// The per-pixel arithmetic is the rounding-pinned sequence of stp_math.cuh.
#include "stp_kernels.cuh"

#ifndef STP_HIER_SPEC
#define STP_HIER_SPEC 0
#endif
#ifndef STP_HIER_MIDBATCH
#define STP_HIER_MIDBATCH 0
#endif
#ifndef STP_HIER_MINB_FWD  // resident CTAs per SM asked from ptxas (register cap 48 / 64): the kernels are latency bound and
#define STP_HIER_MINB_FWD 5  // occupancy limited by registers; A/B on B200: fwd 3->5 CTAs -22 %, bwd 3->4 CTAs -10 % (5: worse, spills)
#endif
#ifndef STP_HIER_MINB_FWD_CULL  // the 4x4-culling forward variant keeps more state live: 4 CTAs (64 registers) beat 5 there
#define STP_HIER_MINB_FWD_CULL 4
#endif
#ifndef STP_HIER_MINB_BWD
#define STP_HIER_MINB_BWD 4
#endif
#ifndef STP_HIER_FIFO      // defer mid/head work through the per-block FIFO until both blocks of a warp have a group
#define STP_HIER_FIFO 0
#endif
#ifndef STP_HIER_COMPACT   // with 4x4 culling: compact the surviving entries of a batch before ranking them
#define STP_HIER_COMPACT 1
#endif
#ifndef STP_HIER_PREFILTER // with 4x4 culling: bounding-box rejection before the exact contribution test
#define STP_HIER_PREFILTER 0
#endif

namespace stp {

namespace {

constexpr float kFltMax = 3.402823466e+38f;
constexpr int kDeadBit = 0x40000000;  // entry cannot reach any pixel of the 4x4 block (ids are < 2^30)
constexpr int kIdMask = 0x3fffffff;
constexpr int kTailStride = 80;  // 64 entries + 16 pad: the two blocks of a warp live in disjoint banks
constexpr int kFifoGroups = 16;  // capacity of the tail -> mid hand-over queue of a block, in groups of 4 ids
constexpr int kFifoIds = kFifoGroups * 4;
constexpr int kFifoLimit = 8;    // backlog a block may keep while the other block of its warp has nothing to do

template <int MID>
struct HierShared {
    static constexpr int kMidCap = MID - 4;       // resident entries after a pop
    static constexpr int kMidStride = MID - 3;    // odd stride: the 8 quads of a warp hit distinct banks
    float tail_d[16 * kTailStride];
    int tail_id[16 * kTailStride];
    float new_d[16 * 48];  // 32 entries per block, stride 48 (16-bank offset between the blocks of a warp)
    int new_id[16 * 48];
    float mid_d[64 * kMidStride];
    int mid_id[64 * kMidStride];
    int out_id[64 * 4];
    int fifo_id[16 * kFifoIds];  // popped tail groups waiting for the mid stage, per block
    float tail_ray[16 * 3];
    float mid_ray[64 * 3];
    float pix_ray[3 * 256];  // per-pixel view ray, [component][thread]: only read by the few entries that pass the alpha test
};
// backward only: per-pixel constants of the gradient formulas, [k][thread] (k: g0 g1 g2 f0 f1 f2 T_final bg_dot).  Kept in
// shared memory instead of eight registers per thread: they are only touched when a pixel actually blends an entry,
// and the kernel is occupancy-limited by registers.
struct HierSharedBwd {
    float pix_const[8 * 256];
};

struct GaussRec {
    float2 xy;
    float4 co;
    float ic[6];
    float ux, uy, uz;
};

__device__ __forceinline__ void load_inv(const float4* __restrict__ inv, int id, float* ic, float& ux, float& uy, float& uz) {
    const float4 a = __ldg(inv + 3 * id), b = __ldg(inv + 3 * id + 1), c = __ldg(inv + 3 * id + 2);
    ic[0] = a.x; ic[1] = a.y; ic[2] = a.z;
    ic[3] = b.x; ic[4] = b.y; ic[5] = b.z;
    ux = c.x; uy = c.y; uz = c.z;
}

// true if no pixel of the 4x4 block at (bx0,by0) can see alpha >= 1/255 from this Gaussian: real q(d) <= thr implies
// |dx| <= sqrt(2 thr C/det), |dy| <= sqrt(2 thr A/det); thr inflated by 0.1 % + 0.01, half-widths by 0.01 px.
__device__ __forceinline__ bool block_unreachable(float2 xy, float4 co, float bx0, float by0) {
    const float det = co.x * co.z - co.y * co.y;
    if (!(det > 0.0f) || !(co.x > 0.0f) || !(co.z > 0.0f) || !(co.w > 0.0f)) return false;
    const float thr = __logf(255.0f * co.w) * 1.001f + 0.01f;
    if (!(thr > 0.0f)) return false;
    const float s = 2.0f * thr / det;
    const float hx = sqrtf(s * co.z) * 1.0001f + 0.01f, hy = sqrtf(s * co.x) * 1.0001f + 0.01f;
    if (!(hx < 1e9f) || !(hy < 1e9f)) return false;
    return (xy.x + hx < bx0) || (xy.x - hx > bx0 + 3.0f) || (xy.y + hy < by0) || (xy.y - hy > by0 + 3.0f);
}

// per-pixel blending state
template <bool BWD>
struct PixelState;
template <>
struct PixelState<false> {
    float T, C0, C1, C2;
};
template <>
struct PixelState<true> {
    float T, C0, C1, C2;
};

template <int HEAD, int MID, bool CULL, bool BWD>
__global__ void __launch_bounds__(256, (HEAD <= 4 && MID <= 12) ? (BWD ? STP_HIER_MINB_BWD : (CULL ? STP_HIER_MINB_FWD_CULL : STP_HIER_MINB_FWD)) : 1)
render_hier_kernel(Frame f, RenderArgs a, RenderBwdArgs ab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Sh = HierShared<MID>;
    Sh& sh = *reinterpret_cast<Sh*>(smem_raw);
    float* const pc = reinterpret_cast<HierSharedBwd*>(smem_raw + sizeof(Sh))->pix_const + threadIdx.x;  // BWD only

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, hl = lane & 15;
    const uint32_t hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int b = warp * 2 + half;        // 4x4 block inside the tile
    const int q = hl >> 2, p = hl & 3;    // quad inside the block, pixel inside the quad
    const int qg = b * 4 + q;             // quad inside the tile
    const uint32_t qmask = 0xfu << (lane & ~3);
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const int cx = tile_x * 16 + (b & 3) * 4, cy = tile_y * 16 + (b >> 2) * 4;
    const int px = cx + (q & 1) * 2 + (p & 1), py = cy + (q >> 1) * 2 + (p >> 1);
    const bool inside = px < f.W && py < f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const size_t plane = (size_t)f.W * f.H;
    const int tile_lin = tile_y * f.grid_x + tile_x;
    if constexpr (BWD) {
        // with a blend log the replay kernel has already handled every pixel whose log is complete: this kernel only
        // re-sorts tiles that contain a pixel with more blends than the log holds, and only for those pixels
        if (ab.blend_rec != nullptr && ab.tile_flags[tile_lin] == 0u) return;
    }

    const uint2* __restrict__ ranges = BWD ? ab.ranges : a.ranges;
    const uint32_t* __restrict__ point_list = BWD ? ab.point_list : a.point_list;
    const float2* __restrict__ means2D = BWD ? ab.means2D : a.means2D;
    const float4* __restrict__ conic_opacity = BWD ? ab.conic_opacity : a.conic_opacity;
    const float4* __restrict__ cov3D_inv = BWD ? ab.cov3D_inv : a.cov3D_inv;
    const float* __restrict__ colors = BWD ? ab.colors : a.colors;

    const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
    {
        const Vec3 r = view_ray(cam, pxf, pyf);  // the pixel's own ray (no half-pixel offset, :355)
        sh.pix_ray[tid] = r.x; sh.pix_ray[256 + tid] = r.y; sh.pix_ray[512 + tid] = r.z;
    }
    if (hl == 0) {
        const Vec3 r = view_ray(cam, fadd((float)cx, 1.5f), fadd((float)cy, 1.5f));
        sh.tail_ray[b * 3 + 0] = r.x; sh.tail_ray[b * 3 + 1] = r.y; sh.tail_ray[b * 3 + 2] = r.z;
    }
    if (p == 0) {
        const Vec3 r = view_ray(cam, fadd((float)cx, 0.5f + 2.0f * (q & 1)), fadd((float)cy, 0.5f + 2.0f * (q >> 1)));
        sh.mid_ray[qg * 3 + 0] = r.x; sh.mid_ray[qg * 3 + 1] = r.y; sh.mid_ray[qg * 3 + 2] = r.z;
    }
    __syncwarp();

    PixelState<BWD> ps;
    ps.T = 1.0f;
    ps.C0 = ps.C1 = ps.C2 = 0.f;
    if constexpr (BWD) {
        float T_final = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, f0 = 0.f, f1 = 0.f, f2 = 0.f;
        if (inside) {
            T_final = ab.final_T[pix_id];
            g0 = ab.dL_dpix[pix_id];
            g1 = ab.dL_dpix[plane + pix_id];
            g2 = ab.dL_dpix[2 * plane + pix_id];
            f0 = ab.pixel_colors[pix_id] - T_final * f.background[0];
            f1 = ab.pixel_colors[plane + pix_id] - T_final * f.background[1];
            f2 = ab.pixel_colors[2 * plane + pix_id] - T_final * f.background[2];
        }
        pc[0] = g0; pc[256] = g1; pc[512] = g2;
        pc[768] = f0; pc[1024] = f1; pc[1280] = f2;
        pc[1536] = T_final;
        pc[1792] = f.background[0] * g0 + f.background[1] * g1 + f.background[2] * g2;
    }
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;
    bool active = inside;
    if constexpr (BWD) {
        if (ab.blend_rec != nullptr) active = inside && ab.blend_count[pix_id] > (uint32_t)ab.rec_cap;
    }
    // forward: next slot of this pixel in its blend log (element index into blend_rec, +256 per blend; the host only
    // enables the log when the whole array can be indexed with 32 bits)
    uint32_t rec_idx = (uint32_t)tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)tid;

    // head queue: sorted by depth, hd[0] is the next to blend
    float hd[HEAD], hs[HEAD];
    int hi[HEAD];
#pragma unroll
    for (int k = 0; k < HEAD; ++k) {
        hd[k] = kFltMax;
        hs[k] = 0.f;
        hi[k] = -1;
    }
    int hcount = 0;

    // ---- blend the head minimum (blend_one, :386-417) -----------------------------------------------------------------
    auto blend_one = [&]() {
        --hcount;
        if (!active) return;
        const int id = hi[0];
        if constexpr (!BWD) {
            const float alpha = hs[0];
            const float test_T = fmul(ps.T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                active = false;
                return;
            }
            ps.C0 = ffma(fmul(__ldg(colors + 3 * id + 0), alpha), ps.T, ps.C0);
            ps.C1 = ffma(fmul(__ldg(colors + 3 * id + 1), alpha), ps.T, ps.C1);
            ps.C2 = ffma(fmul(__ldg(colors + 3 * id + 2), alpha), ps.T, ps.C2);
            ps.T = test_T;
            if (a.blend_rec != nullptr) {
                // streaming store: the log is written once and read by the backward pass much later
                const uint32_t rec_end = ((uint32_t)tile_lin + 1u) * (uint32_t)a.rec_cap * 256u;
                if (rec_idx < rec_end) __stcs(a.blend_rec + rec_idx, make_uint2((uint32_t)id, __float_as_uint(alpha)));
                rec_idx += 256u;
            }
        } else {
            const float G = hs[0];
            const float4 co = __ldg(conic_opacity + id);
            const float alpha = fminf(0.99f, fmul(co.w, G));
            const float test_T = fmul(ps.T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                active = false;
                return;
            }
            const float2 xy = __ldg(means2D + id);
            const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
            const float dchannel_dcolor = alpha * ps.T;
            const float c0 = __ldg(colors + 3 * id + 0), c1 = __ldg(colors + 3 * id + 1), c2 = __ldg(colors + 3 * id + 2);
            ps.C0 += c0 * alpha * ps.T;
            ps.C1 += c1 * alpha * ps.T;
            ps.C2 += c2 * alpha * ps.T;
            const float inv_T = 1.0f / test_T;
            const float g0 = pc[0], g1 = pc[256], g2 = pc[512];
            float dL_dalpha = (c0 - (pc[768] - ps.C0) * inv_T) * g0 + (c1 - (pc[1024] - ps.C1) * inv_T) * g1 +
                              (c2 - (pc[1280] - ps.C2) * inv_T) * g2;
            dL_dalpha *= ps.T;
            dL_dalpha += (-pc[1536] / (1.f - alpha)) * pc[1792];
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y;
            const float dG_ddely = -gdy * co.z - gdx * co.y;
            accumulate_grads(ab.grad_accum, id, dchannel_dcolor * g0, dchannel_dcolor * g1, dchannel_dcolor * g2,
                             dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * dx * dL_dG,
                             -0.5f * gdx * dy * dL_dG, -0.5f * gdy * dy * dL_dG, G * dL_dalpha);
            ps.T = test_T;
        }
#pragma unroll
        for (int k = 1; k < HEAD; ++k) {
            hd[k - 1] = hd[k];
            hs[k - 1] = hs[k];
            hi[k - 1] = hi[k];
        }
        hd[HEAD - 1] = kFltMax;
    };

    // ---- an entry arrives at the pixel (front4OneFromMid inner body, :421-536) ---------------------------------------------
    // Split in two so that the four entries a quad receives together can be evaluated with instruction-level
    // parallelism: head_eval is side-effect free (alpha test first, then -- only for survivors -- the 48-byte
    // inverse covariance and the depth on the pixel's ray); head_insert is the sequential queue update.
    // Entries flagged kDeadBit cannot reach any pixel of this 4x4 block (conservative test at the tail stage):
    // they still flow through every queue (they decide WHEN other entries are popped and blended) but are never
    // evaluated per pixel.
    auto head_eval = [&](int id, float& ed, float& es) -> bool {
        if (id < 0 || (id & kDeadBit) || !active) return false;
        const float2 xy = __ldg(means2D + id);
        const float4 co = __ldg(conic_opacity + id);
        const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
        const float power = gaussian_power(dx, dy, co.x, co.y, co.z);
        if (power > 0.0f) return false;
        const float G = expf(power);
        const float alpha = fminf(0.99f, fmul(co.w, G));
        if (alpha < kAlphaThreshold) return false;
        float ic[6], ux, uy, uz;
        load_inv(cov3D_inv, id, ic, ux, uy, uz);
        const Vec3 ray{sh.pix_ray[tid], sh.pix_ray[256 + tid], sh.pix_ray[512 + tid]};
        ed = depth_along_ray(ic, ux, uy, uz, ray);
        es = BWD ? G : alpha;
        return !(ed < 0.0f);
    };
    auto head_insert = [&](bool ok, int id, float ed, float es) {
        if (hcount >= HEAD) blend_one();
        if (!ok || !active) return;
        int ei = id;
#pragma unroll
        for (int k = 0; k < HEAD; ++k) {
            if (ed < hd[k]) {
                const float td = hd[k], ts = hs[k];
                const int ti = hi[k];
                hd[k] = ed; hs[k] = es; hi[k] = ei;
                ed = td; es = ts; ei = ti;
            }
        }
        ++hcount;
    };

    // ---- mid queue of this quad ---------------------------------------------------------------------------------------
    float* const md = sh.mid_d + qg * Sh::kMidStride;
    int* const mi = sh.mid_id + qg * Sh::kMidStride;
    int* const oi = sh.out_id + qg * 4;
    int mcount = 0, mbase = 0;  // resident entries live at md[mbase .. mbase+mcount)

    // the 4 smallest mid entries (already in oi[0..3]) go to the 4 pixels of the quad
    auto quad_to_head = [&]() {
        const bool any = __any_sync(qmask, active);
        if (!any) return;
        int ids[4];
        float ed[4], es[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ids[k] = oi[k];
#if STP_HIER_SPEC
#pragma unroll
        for (int k = 0; k < 4; ++k) ok[k] = head_eval(ids[k], ed[k], es[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) head_insert(ok[k], ids[k], ed[k], es[k]);
#else
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (hcount >= HEAD) blend_one();
            ok[k] = head_eval(ids[k], ed[k], es[k]);
            if (ok[k]) {
                float e_d = ed[k], e_s = es[k];
                int ei = ids[k];
#pragma unroll
                for (int j = 0; j < HEAD; ++j) {
                    if (e_d < hd[j]) {
                        const float td = hd[j], ts = hs[j];
                        const int ti = hi[j];
                        hd[j] = e_d; hs[j] = e_s; hi[j] = ei;
                        e_d = td; e_s = ts; ei = ti;
                    }
                }
                ++hcount;
            }
        }
#endif
    };

    // one group of 4 tail entries (ids g_id[0..3], quad lane p owns entry p) enters the mid queue (:566-677)
    // depth of one tail entry on the quad's ray (side-effect free: the four groups of a tail pop are evaluated
    // together for instruction-level parallelism, then merged one after the other)
    auto mid_depth = [&](int my_id) -> float {
        float my_d = kFltMax;
        if (my_id >= 0) {
            float ic[6], ux, uy, uz;
            load_inv(cov3D_inv, my_id & kIdMask, ic, ux, uy, uz);
            const Vec3 mr{sh.mid_ray[qg * 3], sh.mid_ray[qg * 3 + 1], sh.mid_ray[qg * 3 + 2]};
            my_d = depth_along_ray(ic, ux, uy, uz, mr);
        }
        return my_d;
    };
    auto mid_push_group = [&](int my_id, float my_d) {
        // rank among the 4 new entries, ties by lane (shflRankingLocal, :129-143)
        float nd[4];
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            nd[o] = __shfl_sync(qmask, my_d, o, 4);
            rank += (o != p) && (nd[o] < my_d || (nd[o] == my_d && o < p));
        }
        // final position of my new entry: rank + #resident <= it (resident first on ties)
        int pos_new = rank;
        for (int k = 0; k < mcount; ++k) pos_new += md[mbase + k] <= my_d;
        // final positions of the resident entries I own (k = p, p+4, ...): k + #new < it
        float rd[Sh::kMidCap / 4];
        int ri[Sh::kMidCap / 4], rpos[Sh::kMidCap / 4];
#pragma unroll
        for (int j = 0; j < Sh::kMidCap / 4; ++j) {
            const int k = p + 4 * j;
            rpos[j] = -1;
            if (k < mcount) {
                rd[j] = md[mbase + k];
                ri[j] = mi[mbase + k];
                rpos[j] = k + (nd[0] < rd[j]) + (nd[1] < rd[j]) + (nd[2] < rd[j]) + (nd[3] < rd[j]);
            }
        }
        __syncwarp(qmask);
        const bool pop = mcount + 4 > MID - 4;
        const int shift = pop ? 4 : 0;
        if (pos_new >= shift) {
            md[pos_new - shift] = my_d;
            mi[pos_new - shift] = my_id;
        } else {
            oi[pos_new] = my_id;
        }
#pragma unroll
        for (int j = 0; j < Sh::kMidCap / 4; ++j) {
            if (rpos[j] >= shift) {
                md[rpos[j] - shift] = rd[j];
                mi[rpos[j] - shift] = ri[j];
            } else if (rpos[j] >= 0) {
                oi[rpos[j]] = ri[j];
            }
        }
        mbase = 0;
        mcount = mcount + 4 - shift;
        __syncwarp(qmask);
        if (pop) quad_to_head();
        __syncwarp(qmask);
    };

    // ---- tail queue of this block -------------------------------------------------------------------------------------
    float* const td = sh.tail_d + b * kTailStride;
    int* const ti = sh.tail_id + b * kTailStride;
    float* const nwd = sh.new_d + b * 48;
    int* const nwi = sh.new_id + b * 48;
    int tcount = 0, tbase = 0;  // resident entries live at td[tbase .. tbase+tcount)

    // tail -> mid hand-over queue of this block (ids only, groups of 4); fhead / fcount are identical in the 16 lanes
    int* const ff = sh.fifo_id + b * kFifoIds;
    int fhead = 0, fcount = 0;
    // process queued groups while both blocks of the warp have one (convergent), or while a backlog exceeds `limit`
    auto consume = [&](int limit) {
        while (true) {
            if (!__any_sync(hmask, active)) fcount = 0;  // nothing downstream of this block is listening any more
            const uint32_t hv = __ballot_sync(0xffffffffu, fcount > 0);
            if (hv == 0) break;
            const bool both = (hv & 0xffffu) != 0 && (hv >> 16) != 0;
            if (!both && !__any_sync(0xffffffffu, fcount > limit)) break;
            if (fcount > 0) {
                const int id_g = ff[(fhead * 4 + p) & (kFifoIds - 1)];
                mid_push_group(id_g, mid_depth(id_g));
                fhead = (fhead + 1) & (kFifoGroups - 1);
                --fcount;
            }
            __syncwarp();
        }
    };

    const uint2 range = ranges[tile_y * f.grid_x + tile_x];
    for (uint32_t progress = range.x; progress < range.y; progress += 32) {
        if (!__any_sync(0xffffffffu, active)) break;  // per-warp early exit (:692)

        // each lane of the half-warp evaluates 2 of the 32 new entries on the block-centre ray
        float e_d[2];
        int e_id[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint32_t src = progress + hl + 16 * s;
            int id = -1;
            float d = kFltMax;
            if (src < range.y) id = (int)__ldg(point_list + src);
            if (id >= 0) {
                const float2 xy = __ldg(means2D + id);
                const float4 co = __ldg(conic_opacity + id);
                // conservative "cannot reach this block" test (bounding box of the alpha >= 1/255 ellipse, inflated far
                // beyond the rounding error of the exact evaluations below and per pixel)
                bool unreachable = false;
                if constexpr (!CULL || STP_HIER_PREFILTER) unreachable = block_unreachable(xy, co, (float)cx, (float)cy);
                bool culled = false;
                if constexpr (CULL) {  // :723-743; an unreachable entry is always rejected by the exact test as well
                    culled = unreachable;
                    if (!culled) {
                        float mx, my;
                        const float pw = max_contrib_power<3, 3>(co.x, co.y, co.z, xy.x, xy.y, (float)cx, (float)cy,
                                                                 fadd((float)cx, 3.0f), fadd((float)cy, 3.0f), mx, my);
                        culled = fminf(0.99f, fmul(co.w, expf(-pw))) < kAlphaThreshold;
                    }
                }
                if (!culled) {
                    float ic[6], ux, uy, uz;
                    load_inv(cov3D_inv, id, ic, ux, uy, uz);
                    const Vec3 tray{sh.tail_ray[b * 3], sh.tail_ray[b * 3 + 1], sh.tail_ray[b * 3 + 2]};
                    d = depth_along_ray(ic, ux, uy, uz, tray);
                    if (!CULL && unreachable) id |= kDeadBit;
                } else {
                    id = -1;
                }
            }
            e_d[s] = d;
            e_id[s] = (d == kFltMax) ? -1 : id;
        }
        // with 4x4 culling only a few of the 32 entries survive: compact them (list order kept) into new_d/new_id so
        // that every later step costs O(valid) instead of O(32)
        constexpr bool COMPACT = CULL && STP_HIER_COMPACT;
        const uint32_t vm0 = __ballot_sync(hmask, e_id[0] >= 0) >> (half * 16);
        const uint32_t vm1 = __ballot_sync(hmask, e_id[1] >= 0) >> (half * 16);
        const int n0 = __popc(vm0), n_valid = n0 + __popc(vm1);
        int cpos[2] = {hl, hl + 16};
        if constexpr (COMPACT) {
            cpos[0] = __popc(vm0 & ((1u << hl) - 1u));
            cpos[1] = n0 + __popc(vm1 & ((1u << hl) - 1u));
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!COMPACT || e_id[s] >= 0) {
                nwd[cpos[s]] = e_d[s];
                nwi[cpos[s]] = e_id[s];
            }
        }
        // my resident entries (positions hl, hl+16 of the resident run) before anything is overwritten
        float r_d[2];
        int r_id[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int k = hl + 16 * s;
            r_d[s] = (k < tcount) ? td[tbase + k] : kFltMax;
            r_id[s] = (k < tcount) ? ti[tbase + k] : -1;
        }
        __syncwarp(hmask);
        // ranks: new entry i -> #new before it (depth, then list position) + #resident <= it;
        //        resident k -> k + #new < it
        int rk_new[2] = {0, 0}, sh_res[2] = {0, 0};
        if constexpr (COMPACT) {
#pragma unroll 4
            for (int j = 0; j < n_valid; ++j) {
                const float dj = nwd[j];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    rk_new[s] += (dj < e_d[s]) || (dj == e_d[s] && j < cpos[s]);
                    sh_res[s] += dj < r_d[s];
                }
            }
        } else {
            // my entries sit at list positions hl (< 16) and hl + 16: against the other half of the batch the
            // position tie-break is decided, so half of the comparisons are a single '<' or '<='
#pragma unroll 8
            for (int j = 0; j < 16; ++j) {
                const float dj = nwd[j];
                rk_new[0] += (dj < e_d[0]) || (dj == e_d[0] && j < hl);
                rk_new[1] += dj <= e_d[1];
                sh_res[0] += dj < r_d[0];
                sh_res[1] += dj < r_d[1];
            }
#pragma unroll 8
            for (int j = 16; j < 32; ++j) {
                const float dj = nwd[j];
                rk_new[0] += dj < e_d[0];
                rk_new[1] += (dj < e_d[1]) || (dj == e_d[1] && (j - 16) < hl);
                sh_res[0] += dj < r_d[0];
                sh_res[1] += dj < r_d[1];
            }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (e_id[s] >= 0) {
                // binary search: number of resident entries with depth <= e_d[s]
                int lo = 0, hi_ = tcount;
                while (lo < hi_) {
                    const int mid = (lo + hi_) >> 1;
                    if (td[tbase + mid] <= e_d[s]) lo = mid + 1; else hi_ = mid;
                }
                rk_new[s] += lo;
            }
        }
        __syncwarp(hmask);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (e_id[s] >= 0) {
                td[rk_new[s]] = e_d[s];
                ti[rk_new[s]] = e_id[s];
            }
            const int k = hl + 16 * s;
            if (k < tcount) {
                td[k + sh_res[s]] = r_d[s];
                ti[k + sh_res[s]] = r_id[s];
            }
        }
        tbase = 0;
        tcount += n_valid;
        __syncwarp(hmask);

        // pop the 16 smallest while more than 32 are held (at most twice, :827-846).  The popped ids are handed to the
        // mid stage through a per-block FIFO: which entries leave the tail, and in which order, does not depend on
        // anything downstream, so the mid / head work of the two blocks of this warp can be deferred until BOTH have a
        // group to process and then runs convergently (with 4x4 culling the two tails fill at different times; pushing
        // each pop through mid and head at once would leave the other half-warp idle).
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            if (tcount > 32) {
#if STP_HIER_FIFO
                ff[((fhead + fcount) * 4 + hl) & (kFifoIds - 1)] = ti[tbase + hl];
                fcount += 4;
#else
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const int id_g = ti[tbase + 4 * g + p];
                    mid_push_group(id_g, mid_depth(id_g));
                }
#endif
                tbase += 16;
                tcount -= 16;
            }
        }
#if STP_HIER_FIFO
        __syncwarp();
        consume(kFifoLimit);
#endif
    }

    // ---- drain: tail -> mid -> head (:855-925) ---------------------------------------------------------------------------
    consume(-1);
    const bool half_alive = __any_sync(hmask, active);
    if (!half_alive) tcount = 0;
    while (__any_sync(0xffffffffu, tcount > 0)) {
        if (tcount > 0) {
            const int did = p < tcount ? ti[tbase + p] : -1;
            mid_push_group(did, mid_depth(did));
            tbase += 4;
            tcount -= min(tcount, 4);
        }
    }
    if (!half_alive) mcount = 0;
    while (__any_sync(0xffffffffu, mcount > 0)) {
        if (mcount > 0) {
            __syncwarp(qmask);
            const int v = mi[mbase + p];
            __syncwarp(qmask);
            oi[p] = v;
            __syncwarp(qmask);
            mbase += 4;
            mcount -= 4;
            quad_to_head();
        }
    }
    while (active && hcount > 0) blend_one();

    if constexpr (!BWD) {
        if (inside) {
            a.final_T[pix_id] = ps.T;
            a.out_color[pix_id] = ffma(ps.T, f.background[0], ps.C0);
            a.out_color[plane + pix_id] = ffma(ps.T, f.background[1], ps.C1);
            a.out_color[2 * plane + pix_id] = ffma(ps.T, f.background[2], ps.C2);
            if (a.blend_rec != nullptr) {
                const uint32_t nrec = (rec_idx - ((uint32_t)tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)tid)) >> 8;
                a.blend_count[pix_id] = nrec;
                if (nrec > (uint32_t)a.rec_cap) atomicOr(a.tile_flags + tile_lin, 1u);
            }
        }
    }
}

// ---- backward by replay (GLOBAL and HIER): every pixel walks its own blend log (front to back, the formulation of the
// reference's hierarchical / k-buffer backward, hierarchical_render.cuh:1071-1170, resorted_render.cuh:303) and
// accumulates the gradients of the logged Gaussians.  Same arithmetic as the BWD branch of blend_one
// above; G is recovered from the logged alpha (alpha / opacity; re-evaluated with expf when alpha was clamped to 0.99).
template <int PIXEL_MAP, bool TILE_FALLBACK>
__global__ void __launch_bounds__(256)
blend_replay_bwd_kernel(Frame f, RenderBwdArgs a) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    int px, py;
    if constexpr (PIXEL_MAP == 2) {  // row-major inside the tile (render_full_kernel)
        px = tile_x * 16 + (tid & 15);
        py = tile_y * 16 + (tid >> 4);
    } else if constexpr (PIXEL_MAP == 1) {  // thread -> pixel map of render_hier_kernel
        const int half = lane >> 4, hl = lane & 15;
        const int b = warp * 2 + half, q = hl >> 2, p = hl & 3;
        px = tile_x * 16 + (b & 3) * 4 + (q & 1) * 2 + (p & 1);
        py = tile_y * 16 + (b >> 2) * 4 + (q >> 1) * 2 + (p >> 1);
    } else {                   // ... of render_global_fwd_kernel
        px = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
        py = tile_y * 16 + (warp >> 1) * 4 + (lane >> 3);
    }
    if (px >= f.W || py >= f.H) return;
    if constexpr (TILE_FALLBACK) {
        if (a.tile_flags[tile_y * f.grid_x + tile_x] != 0u) return;  // the list-driven kernel does this whole tile
    }
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const uint32_t n = a.blend_count[pix_id];
    if (n == 0u || n > (uint32_t)a.rec_cap) return;
    const float pxf = (float)px, pyf = (float)py;
    const size_t plane = (size_t)f.W * f.H;
    const int tile_lin = tile_y * f.grid_x + tile_x;
    const float T_final = a.final_T[pix_id];
    const float g0 = a.dL_dpix[pix_id], g1 = a.dL_dpix[plane + pix_id], g2 = a.dL_dpix[2 * plane + pix_id];
    const float f0 = a.pixel_colors[pix_id] - T_final * f.background[0];
    const float f1 = a.pixel_colors[plane + pix_id] - T_final * f.background[1];
    const float f2 = a.pixel_colors[2 * plane + pix_id] - T_final * f.background[2];
    const float bg_dot = f.background[0] * g0 + f.background[1] * g1 + f.background[2] * g2;
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;
    const uint2* __restrict__ rec = a.blend_rec + (size_t)tile_lin * a.rec_cap * 256 + tid;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint2 nxt = __ldcs(rec);
    for (uint32_t k = 0; k < n; ++k) {
        const uint2 cur = nxt;
        if (k + 1 < n) nxt = __ldcs(rec + (size_t)(k + 1) * 256);
        const int id = (int)cur.x;
        const float alpha = __uint_as_float(cur.y);
        const float4 co = __ldg(a.conic_opacity + id);
        const float2 xy = __ldg(a.means2D + id);
        const float c0 = __ldg(a.colors + 3 * id + 0), c1 = __ldg(a.colors + 3 * id + 1), c2 = __ldg(a.colors + 3 * id + 2);
        const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
        const float G = (alpha < 0.99f) ? alpha / co.w : expf(gaussian_power(dx, dy, co.x, co.y, co.z));
        const float test_T = fmul(T, fsub(1.0f, alpha));
        const float dchannel_dcolor = alpha * T;
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
        accumulate_grads(a.grad_accum, id, dchannel_dcolor * g0, dchannel_dcolor * g1, dchannel_dcolor * g2,
                         dL_dG * dG_ddelx * ddelx_dx, dL_dG * dG_ddely * ddely_dy, -0.5f * gdx * dx * dL_dG,
                         -0.5f * gdx * dy * dL_dG, -0.5f * gdy * dy * dL_dG, G * dL_dalpha);
        T = test_T;
    }
}

template <int HEAD, int MID, bool BWD>
cudaError_t launch_variant(const Frame& f, bool cull, const RenderArgs& a, const RenderBwdArgs& ab, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    const size_t smem = sizeof(HierShared<MID>) + (BWD ? sizeof(HierSharedBwd) : 0);
    if (cull) {
        cudaFuncSetAttribute(render_hier_kernel<HEAD, MID, true, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        render_hier_kernel<HEAD, MID, true, BWD><<<grid, 256, smem, stream>>>(f, a, ab);
    } else {
        cudaFuncSetAttribute(render_hier_kernel<HEAD, MID, false, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        render_hier_kernel<HEAD, MID, false, BWD><<<grid, 256, smem, stream>>>(f, a, ab);
    }
    return cudaGetLastError();
}

template <bool BWD>
cudaError_t dispatch(const Frame& f, const Settings& s, const RenderArgs& a, const RenderBwdArgs& ab, cudaStream_t stream) {
    // instantiated queue sizes: forward.cu:445-488 (HEAD 4/8/16 x MID 8/12/20), backward.cu:739-767 (+HEAD 12)
#define STP_HIER_MID(HEAD_)                                                                   \
    switch (s.q_mid) {                                                                        \
        case 8: return launch_variant<HEAD_, 8, BWD>(f, s.hier_culling, a, ab, stream);       \
        case 12: return launch_variant<HEAD_, 12, BWD>(f, s.hier_culling, a, ab, stream);     \
        case 20: return launch_variant<HEAD_, 20, BWD>(f, s.hier_culling, a, ab, stream);     \
        default: return cudaErrorInvalidValue;                                                \
    }
    switch (s.q_head) {
        case 4: STP_HIER_MID(4)
        case 8: STP_HIER_MID(8)
        case 12:
            if constexpr (BWD) { STP_HIER_MID(12) } else { return cudaErrorInvalidValue; }
        case 16: STP_HIER_MID(16)
        default: return cudaErrorInvalidValue;
    }
#undef STP_HIER_MID
}

}  // namespace

cudaError_t launch_render_hier_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream) {
    RenderBwdArgs dummy{};
    if (a.blend_rec != nullptr) {
        cudaError_t e = cudaMemsetAsync(a.tile_flags, 0, sizeof(uint32_t) * (size_t)f.grid_x * f.grid_y, stream);
        if (e != cudaSuccess) return e;
    }
    return dispatch<false>(f, s, a, dummy, stream);
}

cudaError_t launch_blend_replay_bwd(const Frame& f, const RenderBwdArgs& a, int pixel_map, bool whole_tile_fallback,
                                    cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    if (pixel_map == 1) {
        if (whole_tile_fallback) blend_replay_bwd_kernel<1, true><<<grid, 256, 0, stream>>>(f, a);
        else blend_replay_bwd_kernel<1, false><<<grid, 256, 0, stream>>>(f, a);
    } else if (pixel_map == 0) {
        if (whole_tile_fallback) blend_replay_bwd_kernel<0, true><<<grid, 256, 0, stream>>>(f, a);
        else blend_replay_bwd_kernel<0, false><<<grid, 256, 0, stream>>>(f, a);
    } else {
        blend_replay_bwd_kernel<2, false><<<grid, 256, 0, stream>>>(f, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_render_hier_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream) {
    RenderArgs dummy{};
    if (a.blend_rec != nullptr) {
        cudaError_t e = launch_blend_replay_bwd(f, a, 1, false, stream);
        if (e != cudaSuccess) return e;
    }
    return dispatch<true>(f, s, dummy, a, stream);  // re-sorting backward: everything, or only the pixels whose log overflowed
}

}  // namespace stp
