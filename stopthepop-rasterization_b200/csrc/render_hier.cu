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
//
// How it is done is new.
//   Data movement.  The tile's Gaussians are not gathered through their ids: the tile-sort epilogue has
//   written one 64-byte record per instance in list order (stp_slab.cuh).  The tail stage reads them from
//   a shared-memory ring that is filled 32 records (2 KB) at a time with bulk-async copies (TMA,
//   cp.async.bulk + mbarrier: "full" barriers armed with the byte count, "empty" barriers on which every
//   reading thread arrives), shared by the eight warps of the tile; the queues carry tile-local list
//   positions, and the mid / head stages read "their" record from the (L1/L2-resident) slab.
//   Queues.  One warp owns two 4x4 blocks (one per half-warp); all queue manipulation is rank based
//   (every lane computes the final position of "its" entries with branch-free counting / binary search
//   and scatters them once) instead of compare-exchange networks with a barrier per stage; queues are
//   addressed through base offsets so popping never moves data.
//   Head stage.  A pixel's result depends only on the ORDER in which its quad hands it entries, not on
//   when: the pop "if the head is full" that precedes every arrival only matters before an entry that
//   is actually inserted.  So the mid stage does not call the pixels; it appends the popped positions
//   to a per-quad ring, and the pixels consume that stream at their own pace: every lane scans forward
//   (cheap alpha test) until it holds a survivor, and the expensive part -- depth on the pixel's ray,
//   head insertion, blending of the popped minimum -- runs for the whole warp once most lanes hold one.
//   With the lock-step version only the few pixels that pass the alpha test did useful work in that
//   part, and the two half-warps of a warp (different blocks, different pop times) never overlapped.
// The per-pixel arithmetic is the rounding-pinned sequence of stp_math.cuh.
#include "stp_kernels.cuh"
#include "stp_slab.cuh"

#ifndef STP_HIER_MINB_FWD  // resident CTAs per SM asked from ptxas: the kernels are latency / issue bound and
#define STP_HIER_MINB_FWD 4  // occupancy limited by registers
#endif
#ifndef STP_HIER_MINB_FWD_CULL
#define STP_HIER_MINB_FWD_CULL 4
#endif
#ifndef STP_HIER_MINB_BWD
#define STP_HIER_MINB_BWD 3
#endif
#ifndef STP_HIER_THRESH    // lanes that must hold a survivor before the warp runs the insertion / blending step
#define STP_HIER_THRESH 20
#endif
#ifndef STP_HIER_SCAN      // ring entries a lane may scan between two warp votes
#define STP_HIER_SCAN 2
#endif
#ifndef STP_HIER_RING      // per-quad stream ring (entries, power of two >= 64: one batch can append 32)
#define STP_HIER_RING 64
#endif
#ifndef STP_HIER_COMPACT   // with 4x4 culling: compact the surviving entries of a batch before ranking them
#define STP_HIER_COMPACT 1
#endif

namespace stp {

namespace {

constexpr float kFltMax = 3.402823466e+38f;
constexpr int kDeadBit = 0x40000000;  // entry cannot reach any pixel of the 4x4 block (list positions are < 2^30)
constexpr int kIdxMask = 0x3fffffff;
constexpr float kPowerReject = -5.6f;  // ln(1/255) = -5.541: below this, opacity * exp(power) < 1/255 for any opacity <= 1
constexpr int kTailStride = 80;  // 64 entries + 16 pad: the two blocks of a warp live in disjoint banks
constexpr int kStages = 4;       // slab ring: batches of 32 records in flight / being read
constexpr int kBatch = 32;
constexpr int kRing = STP_HIER_RING;  // per-quad stream ring (entries); one batch can append at most 32
constexpr int kWarps = 8;

template <int MID>
struct HierShared {
    static constexpr int kMidCap = MID - 4;       // resident entries after a pop
    static constexpr int kMidStride = MID - 3;    // odd stride: the 8 quads of a warp hit distinct banks
    float4 slab[kStages][kBatch * kSlabChunks];   // TMA destination, 2 KB per stage
    uint64_t full[kStages];                       // mbarrier per stage: the batch has landed (armed with expect_tx)
    uint64_t empty[kStages];                      // mbarrier per stage: all eight warps have read the batch
    uint32_t released[kStages];                   // elects the warp that arrives last (it refills the stage)
    float tail_d[16 * kTailStride];
    int tail_id[16 * kTailStride];
    float new_d[16 * 48];  // 32 entries per block, stride 48 (16-bank offset between the blocks of a warp)
    int new_id[16 * 48];
    float mid_d[64 * kMidStride];
    int mid_id[64 * kMidStride];
    int ring[64 * kRing];  // per quad: list positions popped by the mid queue, in order
    int popped[64 * 4];    // per quad: the group that is leaving the mid queue
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

template <int HEAD, int MID, bool CULL, bool BWD>
__global__ void __launch_bounds__(256, (HEAD <= 4 && MID <= 12) ? (BWD ? STP_HIER_MINB_BWD : (CULL ? STP_HIER_MINB_FWD_CULL : STP_HIER_MINB_FWD)) : 1)
render_hier_kernel(Frame f, RenderArgs a, RenderBwdArgs ab) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
    } else {
        if (a.abort_flag != nullptr && *a.abort_flag != 0u) return;  // asynchronous forward: binning arena too small
    }

    const uint2* __restrict__ ranges = BWD ? ab.ranges : a.ranges;
    const float4* __restrict__ slab = BWD ? ab.slab : a.slab;
    const float4* __restrict__ slab_rgb = BWD ? ab.slab_rgb : a.slab_rgb;

    const uint2 range = ranges[tile_lin];
    const uint32_t first = range.x;                      // absolute list position of the tile's first instance
    const int n = (int)(range.y - range.x);
    const int nb = (n + kBatch - 1) / kBatch;            // batches of 32 list entries

    // ---- slab ring: the first kStages batches are requested before anything else happens ------------------------------
    auto request_batch = [&](int k) {  // one thread
        const int s = k % kStages;
        const uint32_t bytes = (uint32_t)min(kBatch, n - k * kBatch) * kSlabRecordBytes;
        mbar_expect_tx(&sh.full[s], bytes);
        bulk_g2s(sh.slab[s], slab + (size_t)kSlabChunks * (first + (uint32_t)k * kBatch), bytes, &sh.full[s]);
    };
    // A warp is done reading batch k: it arrives on the stage's "empty" mbarrier (release); the warp that arrives last --
    // elected through a counter -- waits for that phase to complete (acquire: every warp's reads of the stage have been
    // performed), then re-arms the "full" barrier and issues the bulk copy of batch k + kStages into the stage.
    auto release_stage = [&](int k) {
        const int s = k % kStages;
        mbar_arrive(&sh.empty[s]);  // every thread that read the stage arrives itself (256 arrivals per phase)
        __syncwarp();
        if (lane == 0) {
            if (atomicAdd(&sh.released[s], 1u) == kWarps - 1) {
                sh.released[s] = 0u;
                mbar_wait(&sh.empty[s], (uint32_t)(k / kStages) & 1u);
                if (k + kStages < nb) request_batch(k + kStages);
            }
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kWarps * 32);
            sh.released[s] = 0u;
        }
        mbar_fence_init();
        for (int k = 0; k < min(nb, kStages); ++k) request_batch(k);
    }

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

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
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
    __syncthreads();  // barriers initialised, rays written (the only CTA barrier before the epilogue)

    // head queue: sorted by depth, hd[0] is the next to blend; hi = tile-local list position (the blend reads colour and
    // Gaussian id from the {r, g, b, id} slab; the blend log stores the position)
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
        const uint32_t j = first + (uint32_t)hi[0];
        if constexpr (!BWD) {
            const float alpha = hs[0];
            const float test_T = fmul(T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                active = false;
                return;
            }
            const float4 cr = __ldg(slab_rgb + j);
            C0 = ffma(fmul(cr.x, alpha), T, C0);
            C1 = ffma(fmul(cr.y, alpha), T, C1);
            C2 = ffma(fmul(cr.z, alpha), T, C2);
            T = test_T;
            if (a.blend_rec != nullptr) {
                // streaming store: the log is written once and read by the backward pass much later
                const uint32_t rec_end = ((uint32_t)tile_lin + 1u) * (uint32_t)a.rec_cap * 256u;
                if (rec_idx < rec_end) __stcs(a.blend_rec + rec_idx, make_uint2((uint32_t)hi[0], __float_as_uint(alpha)));
                rec_idx += 256u;
            }
        } else {
            const float G = hs[0];
            float4 h0, h1;
            slab_ldg_head(slab, j, h0, h1);
            const float4 co = make_float4(h0.z, h0.w, h1.x, h1.y);
            const float alpha = fminf(0.99f, fmul(co.w, G));
            const float test_T = fmul(T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                active = false;
                return;
            }
            const float4 cr = __ldg(slab_rgb + j);
            const int id = __float_as_int(cr.w);
            const float dx = fsub(h0.x, pxf), dy = fsub(h0.y, pyf);
            const float dchannel_dcolor = alpha * T;
            const float c0 = cr.x, c1 = cr.y, c2 = cr.z;
            C0 += c0 * alpha * T;
            C1 += c1 * alpha * T;
            C2 += c2 * alpha * T;
            const float inv_T = 1.0f / test_T;
            const float g0 = pc[0], g1 = pc[256], g2 = pc[512];
            float dL_dalpha = (c0 - (pc[768] - C0) * inv_T) * g0 + (c1 - (pc[1024] - C1) * inv_T) * g1 +
                              (c2 - (pc[1280] - C2) * inv_T) * g2;
            dL_dalpha *= T;
            dL_dalpha += (-pc[1536] / (1.f - alpha)) * pc[1792];
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
        for (int k = 1; k < HEAD; ++k) {
            hd[k - 1] = hd[k];
            hs[k - 1] = hs[k];
            hi[k - 1] = hi[k];
        }
        hd[HEAD - 1] = kFltMax;
    };

    // ---- per-quad stream ring and the pixel's reader state --------------------------------------------------------------
    int* const rq = sh.ring + qg * kRing;
    uint32_t wr = 0;       // entries appended to my quad's ring so far (identical in the 4 lanes of the quad)
    uint32_t cursor = 0;   // entries of it this pixel has looked at
    int held = -1;         // list position of a survivor of the alpha test that waits for the insertion step
    float held_s = 0.f;    // its alpha (forward) / G (backward)
    // free ring entries of my quad: a slot is reusable once every pixel of the quad that still blends has read it
    auto ring_free = [&]() -> int {
        uint32_t behind = active ? wr - cursor : 0u;
        behind = max(behind, __shfl_xor_sync(0xffffffffu, behind, 1));
        behind = max(behind, __shfl_xor_sync(0xffffffffu, behind, 2));
        return kRing - (int)behind;
    };

    // ---- mid queue of this quad ---------------------------------------------------------------------------------------
    float* const md = sh.mid_d + qg * Sh::kMidStride;
    int* const mi = sh.mid_id + qg * Sh::kMidStride;
    int mcount = 0, mbase = 0;  // resident entries live at md[mbase .. mbase+mcount)

    int* const po = sh.popped + qg * 4;
    // po[0..3] (the four entries that just left the mid queue, in order) -> the quad's ring, live entries only
    auto append_popped = [&]() {
        const int e = po[p];
        const bool live = (e & kDeadBit) == 0;
        const uint32_t lm = (__ballot_sync(qmask, live) >> (lane & ~3)) & 0xfu;
        if (live) rq[(wr + (uint32_t)__popc(lm & ((1u << p) - 1u))) & (kRing - 1)] = e;
        wr += (uint32_t)__popc(lm);
        __syncwarp(qmask);
    };
    // depth of one tail entry on the quad's ray
    // depth of one tail entry on the quad's ray.  (Flagging, here, the entries whose alpha >= 1/255 ellipse cannot reach
    // the 2x2 quad so that its pixels skip them unevaluated was measured 9 % SLOWER at C3b: the test costs more than
    // the scans it saves.)
    auto mid_depth = [&](int my_e) -> float {
        float my_d = kFltMax;
        if (my_e >= 0) {
            const uint32_t j = first + (uint32_t)(my_e & kIdxMask);
            const float4* const rec = slab + 4 * (size_t)j;
            const uint32_t sw = slab_swizzle(j);
            const float4 c1 = __ldg(rec + (1 ^ sw)), c2 = __ldg(rec + (2 ^ sw)), c3 = __ldg(rec + (3 ^ sw));
            const float ic[6] = {c1.w, c2.x, c2.y, c2.z, c2.w, c3.x};
            const Vec3 mr{sh.mid_ray[qg * 3], sh.mid_ray[qg * 3 + 1], sh.mid_ray[qg * 3 + 2]};
            my_d = depth_along_ray(ic, c3.y, c3.z, c3.w, mr);
        }
        return my_d;
    };
    // one group of 4 tail entries (quad lane p owns entry p) enters the mid queue (:566-677); the 4 smallest leave for the
    // quad's stream ring when the queue would hold more than MID-4
    auto mid_push_group = [&](int my_e, float my_d) {
        // rank among the 4 new entries, ties by lane (shflRankingLocal, :129-143)
        float nd[4];
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            nd[o] = __shfl_sync(qmask, my_d, o, 4);
            rank += (o < p) ? (nd[o] <= my_d) : (nd[o] < my_d);  // ties by position; o == p: my own entry, never counted
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
        // the popped four, by rank, pass through `po` so that entries no pixel of the block will ever evaluate (dead bit,
        // padding) can be left out of the stream: they had to flow through the queues, the pixels need not see them
        // (with 4x4 culling nothing dead ever enters the queues: the popped four go straight to the ring)
        int* const out = CULL ? rq : po;
        const uint32_t obase = CULL ? wr : 0u;
        constexpr uint32_t omask = CULL ? (uint32_t)(kRing - 1) : 3u;
        if (pos_new >= shift) {
            md[pos_new - shift] = my_d;
            mi[pos_new - shift] = my_e;
        } else {
            out[(obase + (uint32_t)pos_new) & omask] = my_e;
        }
#pragma unroll
        for (int j = 0; j < Sh::kMidCap / 4; ++j) {
            if (rpos[j] >= shift) {
                md[rpos[j] - shift] = rd[j];
                mi[rpos[j] - shift] = ri[j];
            } else if (rpos[j] >= 0) {
                out[(obase + (uint32_t)rpos[j]) & omask] = ri[j];
            }
        }
        mbase = 0;
        mcount = mcount + 4 - shift;
        if constexpr (CULL) wr += (uint32_t)shift;
        __syncwarp(qmask);
        if constexpr (!CULL) {
            if (pop) append_popped();
        }
    };

    // ---- tail queue of this block -------------------------------------------------------------------------------------
    float* const td = sh.tail_d + b * kTailStride;
    int* const ti = sh.tail_id + b * kTailStride;
    float* const nwd = sh.new_d + b * 48;
    int* const nwi = sh.new_id + b * 48;
    int tcount = 0, tbase = 0;  // resident entries live at td[tbase .. tbase+tcount)
    const Vec3 tray{sh.tail_ray[b * 3], sh.tail_ray[b * 3 + 1], sh.tail_ray[b * 3 + 2]};

    // batch k of the tile list: 32 records from the slab ring -> tail queue of both blocks of this warp -> pops -> mid
    auto tail_batch = [&](int k, bool block_alive) {
        const int s = k % kStages;
        mbar_wait(&sh.full[s], (uint32_t)(k / kStages) & 1u);
        const float4* const stage = sh.slab[s];
        // each lane of the half-warp evaluates 2 of the 32 new entries on the block-centre ray
        float e_d[2];
        int e_id[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int slot = hl + 16 * t;
            const int src = k * kBatch + slot;  // tile-local list position
            int e = -1;
            float d = kFltMax;
            if (src < n && block_alive) {
                const uint32_t j = first + (uint32_t)src;
                float4 c0, c1;
                slab_load_head(stage + 4 * slot, j, c0, c1);
                const float2 xy = make_float2(c0.x, c0.y);
                const float4 co = make_float4(c0.z, c0.w, c1.x, c1.y);
                // conservative "cannot reach this block" test (bounding box of the alpha >= 1/255 ellipse, inflated far
                // beyond the rounding error of the exact evaluations below and per pixel)
                bool unreachable = false;
                if constexpr (!CULL) unreachable = block_unreachable(xy, co, (float)cx, (float)cy);
                bool culled = false;
                if constexpr (CULL) {  // :723-743
                    float mx, my;
                    const float pw = max_contrib_power<3, 3>(co.x, co.y, co.z, xy.x, xy.y, (float)cx, (float)cy,
                                                             fadd((float)cx, 3.0f), fadd((float)cy, 3.0f), mx, my);
                    culled = (-pw < kPowerReject && co.w <= 1.0f) || fminf(0.99f, fmul(co.w, expf(-pw))) < kAlphaThreshold;
                }
                if (!culled) {
                    float4 c2, c3;
                    slab_load_tail(stage + 4 * slot, j, c2, c3);
                    const float ic[6] = {c1.w, c2.x, c2.y, c2.z, c2.w, c3.x};
                    d = depth_along_ray(ic, c3.y, c3.z, c3.w, tray);
                    e = (!CULL && unreachable) ? (src | kDeadBit) : src;
                }
            }
            e_d[t] = d;
            e_id[t] = (d == kFltMax) ? -1 : e;
        }
        // this warp is done with the stage: the last of the eight warps re-arms it with batch k + kStages
        release_stage(k);
        // with 4x4 culling only a few of the 32 entries survive: compact them (list order kept) into new_d/new_id so
        // that every later step costs O(valid) instead of O(32)
        constexpr bool COMPACT = CULL && STP_HIER_COMPACT;
        const uint32_t vm0 = __ballot_sync(0xffffffffu, e_id[0] >= 0) >> (half * 16) & 0xffffu;
        const uint32_t vm1 = __ballot_sync(0xffffffffu, e_id[1] >= 0) >> (half * 16) & 0xffffu;
        const int n0 = __popc(vm0), n_valid = n0 + __popc(vm1);
        int cpos[2] = {hl, hl + 16};
        if constexpr (COMPACT) {
            cpos[0] = __popc(vm0 & ((1u << hl) - 1u));
            cpos[1] = n0 + __popc(vm1 & ((1u << hl) - 1u));
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (!COMPACT || e_id[t] >= 0) {
                nwd[cpos[t]] = e_d[t];
                nwi[cpos[t]] = e_id[t];
            }
        }
        // my resident entries (positions hl, hl+16 of the resident run) before anything is overwritten
        float r_d[2];
        int r_id[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int kk = hl + 16 * t;
            r_d[t] = (kk < tcount) ? td[tbase + kk] : kFltMax;
            r_id[t] = (kk < tcount) ? ti[tbase + kk] : -1;
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
                for (int t = 0; t < 2; ++t) {
                    rk_new[t] += (j < cpos[t]) ? (dj <= e_d[t]) : (dj < e_d[t]);  // ties by list position
                    sh_res[t] += dj < r_d[t];
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
        for (int t = 0; t < 2; ++t) {
            if (e_id[t] >= 0) {
                // binary search: number of resident entries with depth <= e_d[t]
                int lo = 0, hi_ = tcount;
                while (lo < hi_) {
                    const int mid = (lo + hi_) >> 1;
                    if (td[tbase + mid] <= e_d[t]) lo = mid + 1; else hi_ = mid;
                }
                rk_new[t] += lo;
            }
        }
        __syncwarp(hmask);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (e_id[t] >= 0) {
                td[rk_new[t]] = e_d[t];
                ti[rk_new[t]] = e_id[t];
            }
            const int kk = hl + 16 * t;
            if (kk < tcount) {
                td[kk + sh_res[t]] = r_d[t];
                ti[kk + sh_res[t]] = r_id[t];
            }
        }
        tbase = 0;
        tcount += n_valid;
        __syncwarp(hmask);

        // pop the 16 smallest while more than 32 are held (at most twice, :827-846): four groups of 4 through the mid stage
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            if (tcount > 32) {
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const int e_g = ti[tbase + 4 * g + p];
                    mid_push_group(e_g, mid_depth(e_g));
                }
                tbase += 16;
                tcount -= 16;
            }
        }
    };

    // ---- the pixel consumes its quad's stream (front4OneFromMid inner body, :421-536) ------------------------------------
    // scan: one ring entry per call, the alpha test only (xy + conic/opacity = first half of the record)
    auto scan_one = [&]() {
        const int e = rq[cursor & (kRing - 1)];
        ++cursor;
        if (e & kDeadBit) return;  // cannot reach this block (or padding, -1): flows through the queues, never evaluated
        float4 c0, c1;
        slab_ldg_head(slab, first + (uint32_t)e, c0, c1);
        const float dx = fsub(c0.x, pxf), dy = fsub(c0.y, pyf);
        const float power = gaussian_power(dx, dy, c0.z, c0.w, c1.x);
        if (power > 0.0f) return;
        // exp(power) < exp(-5.6) < 1/255: with an opacity of at most one the alpha test below fails whatever the rounding
        if (power < kPowerReject && c1.y <= 1.0f) return;
        const float G = expf(power);
        const float alpha = fminf(0.99f, fmul(c1.y, G));
        if (alpha < kAlphaThreshold) return;
        held = e;
        held_s = BWD ? G : alpha;
    };
    // insertion step for the held survivor: depth on the pixel's ray (second half of the record), blend the head minimum
    // if the head is full, insert by strict '<'
    auto insert_held = [&]() {
        int ei = held;
        held = -1;
        float ic[6], ux, uy, uz;
        slab_ldg_inv(slab, first + (uint32_t)ei, ic, ux, uy, uz);
        const Vec3 ray{sh.pix_ray[tid], sh.pix_ray[256 + tid], sh.pix_ray[512 + tid]};
        float e_d = depth_along_ray(ic, ux, uy, uz, ray);
        if (e_d < 0.0f) return;
        if (hcount >= HEAD) blend_one();
        if (!active) return;
        float e_s = held_s;
#pragma unroll
        for (int k = 0; k < HEAD; ++k) {
            if (e_d < hd[k]) {
                const float td_ = hd[k], ts = hs[k];
                const int ti_ = hi[k];
                hd[k] = e_d; hs[k] = e_s; hi[k] = ei;
                e_d = td_; e_s = ts; ei = ti_;
            }
        }
        ++hcount;
    };

    // ---- the warp alternates between producing (tail batches -> mid -> rings) and consuming (pixels read their rings) ----
    int k = 0;             // next batch
    bool stream_end = false;
    while (true) {
        const uint32_t alive = __ballot_sync(0xffffffffu, active);
        if (alive == 0u) break;
        const bool block_alive = ((alive >> (half * 16)) & 0xffffu) != 0u;
        if (!block_alive) {  // nothing downstream of this block listens any more
            tcount = 0;
            mcount = 0;
        }
        // ---- produce --------------------------------------------------------------------------------------------------
        while (true) {
            const int fr = ring_free();
            const uint32_t ok32 = __ballot_sync(0xffffffffu, fr >= 32 || !active);
            const uint32_t ok4 = __ballot_sync(0xffffffffu, fr >= 4 || !active);
            if (k < nb) {
                if (ok32 != 0xffffffffu) break;
                tail_batch(k, block_alive);
                ++k;
                continue;
            }
            // drain: tail -> mid -> rings (:855-925), four entries per step and quad
            const bool room = ((ok4 >> (half * 16)) & 0xffffu) == 0xffffu;
            bool did = false;
            if (room && tcount > 0) {
                const int did_e = p < tcount ? ti[tbase + p] : -1;
                mid_push_group(did_e, mid_depth(did_e));
                tbase += 4;
                tcount -= min(tcount, 4);
                did = true;
            } else if (room && mcount > 0) {
                __syncwarp(qmask);
                po[p] = mi[mbase + p];
                __syncwarp(qmask);
                append_popped();
                mbase += 4;
                mcount -= 4;
                did = true;
            }
            if (!__any_sync(0xffffffffu, did)) break;
        }
        stream_end = __all_sync(0xffffffffu, k >= nb && tcount == 0 && mcount == 0);
        // ---- consume --------------------------------------------------------------------------------------------------
        while (true) {
#pragma unroll 1
            for (int it = 0; it < STP_HIER_SCAN; ++it) {  // several ring entries per vote: most fail the alpha test
                if (active && held < 0 && cursor != wr) scan_one();
            }
            const uint32_t ready = __ballot_sync(0xffffffffu, active && held >= 0);
            const uint32_t scan = __ballot_sync(0xffffffffu, active && held < 0 && cursor != wr);
            if (ready == 0u && scan == 0u) break;
            if (__popc(ready) >= STP_HIER_THRESH) {
                if (active && held >= 0) insert_held();
            } else if (scan == 0u) {
                // nobody can scan any further and only a few lanes hold a survivor: they keep it for the next round unless
                // the stream has ended or the producer needs the ring space their cursors pin
                bool flush = stream_end;
                if (!flush) {
                    const int fr = ring_free();
                    flush = !__all_sync(0xffffffffu, fr >= 32 || !active);
                }
                if (!flush) break;
                if (active && held >= 0) insert_held();
            }
        }
        if (stream_end && !__any_sync(0xffffffffu, active && (held >= 0 || cursor != wr))) break;
    }
    while (active && hcount > 0) blend_one();

    if constexpr (!BWD) {
        if (inside) {
            a.final_T[pix_id] = T;
            a.out_color[pix_id] = ffma(T, f.background[0], C0);
            a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
            a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
            if (a.blend_rec != nullptr) {
                const uint32_t nrec = (rec_idx - ((uint32_t)tile_lin * (uint32_t)a.rec_cap * 256u + (uint32_t)tid)) >> 8;
                a.blend_count[pix_id] = nrec;
                if (nrec > (uint32_t)a.rec_cap) atomicOr(a.tile_flags + tile_lin, 1u);
            }
        }
    }

    // ---- leave the slab ring in order: a warp that stops early still has to release every remaining batch (the other
    // warps' refills wait for all eight), and no bulk copy may be in flight when the CTA exits -----------------------------
    for (; k < nb; ++k) {
        mbar_wait(&sh.full[k % kStages], (uint32_t)(k / kStages) & 1u);
        release_stage(k);
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
    // what the log stores per blend: the tile-local list position in the slab modes (HIER / PPX_FULL: geometry and colour
    // come from the tile's contiguous slabs), the Gaussian id in GLOBAL mode (gathered)
    const uint32_t first = PIXEL_MAP != 0 ? a.ranges[tile_lin].x : 0u;
    uint2 nxt = __ldcs(rec);
    for (uint32_t k = 0; k < n; ++k) {
        const uint2 cur = nxt;
        if (k + 1 < n) nxt = __ldcs(rec + (size_t)(k + 1) * 256);
        const float alpha = __uint_as_float(cur.y);
        int id;
        float4 co;
        float2 xy;
        float c0, c1, c2;
        if constexpr (PIXEL_MAP != 0) {
            const uint32_t j = first + cur.x;
            float4 h0, h1;
            slab_ldg_head(a.slab, j, h0, h1);
            const float4 cr = __ldg(a.slab_rgb + j);
            xy = make_float2(h0.x, h0.y);
            co = make_float4(h0.z, h0.w, h1.x, h1.y);
            c0 = cr.x; c1 = cr.y; c2 = cr.z;
            id = __float_as_int(cr.w);
        } else {
            id = (int)cur.x;
            co = __ldg(a.conic_opacity + id);
            xy = __ldg(a.means2D + id);
            c0 = __ldg(a.colors + 3 * id + 0); c1 = __ldg(a.colors + 3 * id + 1); c2 = __ldg(a.colors + 3 * id + 2);
        }
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
        accumulate_grads(a.grad_accum, a.P, id, dchannel_dcolor * g0, dchannel_dcolor * g1, dchannel_dcolor * g2,
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
