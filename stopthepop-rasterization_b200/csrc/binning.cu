// binning.cu -- tile binning and depth sort: per-tile buckets + one in-shared-memory sort per tile.
//
// Replaces: duplicateWithKeysCUDA (forward.cu:25-65), duplicateWithKeys_extended<TBC,LB,ORDER>
// (stopthepop_common.cuh:324-621), cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:344-352) and
// identifyTileRanges (rasterizer_impl.cu:133-158).
//
// The reference sorts all R (tile << 32 | depth bits) keys with a stable 6-pass LSD radix sort (152 B of HBM
// traffic per instance) and then finds the tile boundaries.  The result of that sort is fully determined:
// instances grouped by tile id, inside a tile ascending by the raw depth bits, ties in emission order =
// ascending Gaussian index.  Here the same order is produced without any global sort pass:
//   1. preprocess.cu histograms the instances per tile (it visits every (Gaussian, tile) pair anyway);
//   2. tile_scan_kernel turns the histogram into bucket offsets -- these ARE the tile ranges;
//   3. duplicate_kernel claims a slot in its tile's bucket per instance (atomic cursor) and stores one 8-byte
//      record (depth bits << 32 | Gaussian index): unique per tile, so any comparison sort reproduces the
//      stable order regardless of the (non-deterministic) slot order;
//   4. tile_sort_* sorts every bucket in shared memory (bitonic network, warp-synchronous below 64-element
//      span) and writes point_list and the sorted 64-bit keys.  Tiles longer than the largest shared-memory
//      capacity are chunk-sorted and merged through a global ping-pong buffer (merge path).
// HBM traffic: 8 B written + 8 B read + 12 B written per instance.
#include "stp_kernels.cuh"
#include "stp_sort.cuh"
#include "stp_slab.cuh"

namespace stp {

namespace {

constexpr int kSeqTiles = 8;

struct DupGaussian {
    float2 xy;
    float4 co;       // conic + opacity (only read if TBC or PTD_MAX)
    float ic[6];     // inverse covariance (PTD only)
    float ux, uy, uz;
    float depth;     // global depth (Z / DISTANCE)
    float thr;
    int x0, y0, w, n;  // rect origin, width, tile count
    uint32_t idx;
};

// evaluates one tile of one Gaussian: returns whether a key is emitted and its depth.
template <bool TBC, int ORDER>
__device__ __forceinline__ bool eval_tile(const DupGaussian& gs, const RayCam& cam, int tx, int ty, uint32_t grid_x,
                                          uint32_t& tile_id, float& depth) {
    tile_id = (uint32_t)ty * grid_x + (uint32_t)tx;
    const float tmin_x = (float)(tx * 16), tmin_y = (float)(ty * 16);
    const float tmax_x = (float)(tx * 16 + 15), tmax_y = (float)(ty * 16 + 15);
    float mx = 0.f, my = 0.f, power = 0.f;
    if constexpr (TBC || ORDER == 3)
        power = max_contrib_power<15, 15>(gs.co.x, gs.co.y, gs.co.z, gs.xy.x, gs.xy.y, tmin_x, tmin_y, tmax_x, tmax_y, mx, my);
    if constexpr (ORDER == 2 || ORDER == 3) {
        float px, py;
        if constexpr (ORDER == 3) {
            px = mx;
            py = my;
        } else {
            px = fmul(fadd(tmin_x, tmax_x), 0.5f);
            py = fmul(fadd(tmin_y, tmax_y), 0.5f);
        }
        const Vec3 d = view_ray(cam, px, py);
        depth = per_tile_depth_key(gs.ic, gs.ux, gs.uy, gs.uz, d);
    } else {
        depth = gs.depth;
    }
    return !TBC || power <= gs.thr;
}

__device__ __forceinline__ void emit_instance(uint32_t* __restrict__ cursor, uint64_t* __restrict__ bucket, uint32_t cap,
                                              uint32_t tile_id, float depth, uint32_t idx) {
    const uint32_t slot = atomicAdd(cursor + tile_id, 1u);
    if (slot < cap) bucket[slot] = ((uint64_t)__float_as_uint(depth) << 32) | (uint64_t)idx;
}

#ifndef STP_DUP_MINB
#define STP_DUP_MINB 4  // 64 registers: A/B on B200 (C3b duplicate 0.50 -> 0.34 ms, C5 1.20 -> 0.77 ms)
#endif
template <bool TBC, int ORDER>
__global__ void __launch_bounds__(256, STP_DUP_MINB)
duplicate_kernel(int P, Frame f, GeometryState g, const int* __restrict__ radii, uint32_t* __restrict__ cursor,
                 uint64_t* __restrict__ bucket, uint32_t cap) {
    if (g.counters[kAbortFlag] != 0u) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    constexpr bool NEED_CO = TBC || ORDER == 3;
    constexpr bool PTD = ORDER == 2 || ORDER == 3;

    RayCam cam;
    if constexpr (PTD) cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);

    DupGaussian gs;
    gs.n = 0;
    gs.w = 1;
    gs.x0 = gs.y0 = 0;
    gs.idx = (uint32_t)idx;
    gs.thr = 0.f;
    gs.depth = 0.f;
    if (idx < P && radii[idx] > 0) {
        gs.xy = g.means2D[idx];
        const float2 ext = g.rects2D[idx];
        const TileRect rc = tile_rect(gs.xy, ext, f.grid_x, f.grid_y, f.row0, f.row1);
        gs.x0 = rc.x0;
        gs.y0 = rc.y0;
        gs.w = max(rc.x1 - rc.x0, 1);
        gs.n = (rc.x1 - rc.x0) * (rc.y1 - rc.y0);
        gs.depth = g.depths[idx];
        if constexpr (NEED_CO) {
            gs.co = g.conic_opacity[idx];
            gs.thr = logf(fdiv(gs.co.w, kAlphaThreshold));
        }
        if constexpr (PTD) {
            const float4 a = g.cov3D_inv[3 * idx], b = g.cov3D_inv[3 * idx + 1], c = g.cov3D_inv[3 * idx + 2];
            gs.ic[0] = a.x; gs.ic[1] = a.y; gs.ic[2] = a.z;
            gs.ic[3] = b.x; gs.ic[4] = b.y; gs.ic[5] = b.z;
            gs.ux = c.x; gs.uy = c.y; gs.uz = c.z;
        }
    }

    // sequential head of the rectangle, four tiles at a time: the four slot claims (atomics with a return value, one
    // L2 round trip each) are issued back to back before the first dependent store, and the tile walk is incremental
    // (no integer division)
    {
        int tx = gs.x0, ty = gs.y0;
        const int n_seq = min(gs.n, kSeqTiles);
        for (int t0 = 0; t0 < n_seq; t0 += 4) {
            uint32_t tile_id[4], slot[4];
            float depth[4];
            bool emit[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                emit[k] = false;
                tile_id[k] = 0;
                depth[k] = 0.f;
                if (t0 + k < n_seq) {
                    emit[k] = eval_tile<TBC, ORDER>(gs, cam, tx, ty, (uint32_t)f.grid_x, tile_id[k], depth[k]);
                    if (++tx == gs.x0 + gs.w) {
                        tx = gs.x0;
                        ++ty;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) slot[k] = emit[k] ? atomicAdd(cursor + tile_id[k], 1u) : 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (emit[k] && slot[k] < cap) bucket[slot[k]] = ((uint64_t)__float_as_uint(depth[k]) << 32) | (uint64_t)gs.idx;
        }
    }

    // warp-cooperative remainder
    uint32_t big = __ballot_sync(0xffffffffu, gs.n > kSeqTiles);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        DupGaussian o;
        o.xy.x = __shfl_sync(0xffffffffu, gs.xy.x, src);
        o.xy.y = __shfl_sync(0xffffffffu, gs.xy.y, src);
        if constexpr (NEED_CO) {
            o.co.x = __shfl_sync(0xffffffffu, gs.co.x, src);
            o.co.y = __shfl_sync(0xffffffffu, gs.co.y, src);
            o.co.z = __shfl_sync(0xffffffffu, gs.co.z, src);
            o.thr = __shfl_sync(0xffffffffu, gs.thr, src);
        }
        if constexpr (PTD) {
#pragma unroll
            for (int k = 0; k < 6; ++k) o.ic[k] = __shfl_sync(0xffffffffu, gs.ic[k], src);
            o.ux = __shfl_sync(0xffffffffu, gs.ux, src);
            o.uy = __shfl_sync(0xffffffffu, gs.uy, src);
            o.uz = __shfl_sync(0xffffffffu, gs.uz, src);
        }
        o.depth = __shfl_sync(0xffffffffu, gs.depth, src);
        o.x0 = __shfl_sync(0xffffffffu, gs.x0, src);
        o.y0 = __shfl_sync(0xffffffffu, gs.y0, src);
        o.w = __shfl_sync(0xffffffffu, gs.w, src);
        o.n = __shfl_sync(0xffffffffu, gs.n, src);
        o.idx = __shfl_sync(0xffffffffu, gs.idx, src);
        for (int t = kSeqTiles + lane; t < o.n; t += 32) {
            uint32_t tile_id = 0;
            float depth = 0.f;
            if (eval_tile<TBC, ORDER>(o, cam, o.x0 + t % o.w, o.y0 + t / o.w, (uint32_t)f.grid_x, tile_id, depth))
                emit_instance(cursor, bucket, cap, tile_id, depth, o.idx);
        }
    }
}

// ---- histogram -> bucket offsets = tile ranges (identifyTileRanges, rasterizer_impl.cu:133-158: tiles without
// instances keep the (0,0) of the reference's memset) --------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kSmallCap = 2048;    // entries sorted by one 256-thread CTA in 16 KB of shared memory
constexpr int kLargeCap = 16384;   // entries sorted by one 1024-thread CTA in 128 KB of shared memory
constexpr int kLargeThreads = 1024;

// CTA b owns tiles [b*1024, (b+1)*1024): it first sums every count before its segment (the whole histogram is a few
// tens of KB in L2, so a redundant read is cheaper than any inter-CTA dependency), then scans its own segment.
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int tiles, const uint32_t* __restrict__ count, uint2* __restrict__ ranges, uint32_t* __restrict__ cursor,
                 uint32_t* __restrict__ large_tiles, uint32_t* __restrict__ counters, uint32_t capacity) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_before;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg0 = blockIdx.x * kScanThreads;
    uint32_t before = 0;
    for (int t = tid; t < seg0; t += kScanThreads) before += count[t];
    const int t = seg0 + tid;
    const uint32_t c = t < tiles ? count[t] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    if (tid == 0) s_before = 0;
    __syncthreads();
    if (lane == 0 && before != 0) atomicAdd(&s_before, before);
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t end = s_before + s_warp[warp] + incl, start = end - c;
    if (t < tiles) {
        ranges[t] = c ? make_uint2(start, end) : make_uint2(0u, 0u);
        cursor[t] = start;
        if (c > (uint32_t)kSmallCap) large_tiles[atomicAdd(counters + 3, 1u)] = (uint32_t)t;
        if (t == tiles - 1) {
            counters[1] = end;  // R
            // asynchronous forward (api.cu): the binning arena was sized before R was known.  If it is too small every
            // later kernel of this frame returns at once (kAbortFlag) and the host re-runs the frame when it looks at R.
            counters[kAbortFlag] = end > capacity ? 1u : 0u;
        }
    }
}

// ---- bitonic sort of one bucket: n unique 64-bit records, virtually padded with ~0 to np (a power of two).
// Every warp keeps a span of 32*E consecutive elements in registers (lane L holds elements L, L+32, ... of the span):
// compare-exchange distances below 32 are warp shuffles, distances 32 .. 16*E are register-to-register, and only
// distances of a whole span or more go through shared memory -- for a 512-entry tile 6 of the 45 stages.
// sorts src[0,n) and hands element i of the sorted sequence to emit(i, value).  s: np * 8 bytes of shared memory.
template <int E, int THREADS, typename Emit>
__device__ __forceinline__ void bitonic_sort_tile(const uint64_t* src, int n, int np, uint64_t* __restrict__ s, int tid,
                                                  Emit emit) {
    constexpr int SPAN = 32 * E;
    const int lane = tid & 31, base = (tid >> 5) * SPAN;
    const bool act = base < np;  // warp-uniform
    uint64_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = base + e * 32 + lane;
        v[e] = (act && i < n) ? src[i] : ~0ull;
    }
    if (act) {
#pragma unroll
        for (int k = 2; k <= SPAN; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) reg_stage<E>(v, j, k, base, lane);
        }
    }
    const int half = np >> 1;
    for (int k = 2 * SPAN; k <= np; k <<= 1) {
        if (act) {
#pragma unroll
            for (int e = 0; e < E; ++e) s[base + e * 32 + lane] = v[e];
        }
        __syncthreads();
        for (int j = k >> 1; j >= SPAN; j >>= 1) {
            for (int t = tid; t < half; t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const uint64_t a = s[i], b = s[l];
                if ((a > b) == ((i & k) == 0)) {
                    s[i] = b;
                    s[l] = a;
                }
            }
            __syncthreads();
        }
        if (act) {
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = s[base + e * 32 + lane];
#pragma unroll
            for (int j = SPAN >> 1; j > 0; j >>= 1) reg_stage<E>(v, j, k, base, lane);
        }
    }
    if (act) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = base + e * 32 + lane;
            if (i < n) emit(i, v[e]);
        }
    }
}

__device__ __forceinline__ int pow2_at_least(int n) {
    int np = 64;
    while (np < n) np <<= 1;
    return np;
}

// what the sort epilogue needs to write the tile's slab (stp_slab.cuh); slab == nullptr: index buffers only
struct SlabSource {
    float4* slab;
    float4* slab_rgb;
    const float2* means2D;
    const float4* conic_opacity;
    const float4* cov3D_inv;
    const float* colors;
};
struct EmitSorted {  // final outputs: the reference's point_list_keys / point_list, and the slab record of the instance
    uint64_t* keys;
    uint32_t* point_list;
    uint64_t tile_hi;
    uint32_t first;  // absolute list position of the tile's first instance
    SlabSource src;
    __device__ __forceinline__ void operator()(int i, uint64_t v) const {
        keys[i] = tile_hi | (v >> 32);
        const uint32_t id = (uint32_t)v;
        point_list[i] = id;
        if (src.slab != nullptr) {
            const float4* const inv = src.cov3D_inv + 3 * (size_t)id;
            slab_store(src.slab, first + (uint32_t)i, (int)id, __ldg(src.means2D + id), __ldg(src.conic_opacity + id),
                       __ldg(inv), __ldg(inv + 1), __ldg(inv + 2));
            const float* const c = src.colors + 3 * (size_t)id;
            src.slab_rgb[first + (uint32_t)i] = make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), __int_as_float((int)id));
        }
    }
};
struct EmitRaw {  // sorted chunk back to the bucket (input of the merge passes)
    uint64_t* dst;
    __device__ __forceinline__ void operator()(int i, uint64_t v) const { dst[i] = v; }
};

// the preprocess histogram and the duplicate kernel's emission disagree for this tile (a bug, or corrupted inputs):
// the bucket holds stale records.  Raised in the geometry arena (counters[2], bit 2) and in the library's mapped host
// word, which the host inspects at its next entry (api.cu: host_error_flags)
__device__ __forceinline__ void flag_binning_mismatch(uint32_t* counters, uint32_t* host_flags) {
    atomicOr(counters + 2, 2u);
    if (host_flags != nullptr) {
        *reinterpret_cast<volatile uint32_t*>(host_flags) = 1u;
        __threadfence_system();
    }
}

// one CTA per tile, tiles of at most kSmallCap instances
__global__ void __launch_bounds__(256)
tile_sort_small_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ cursor,
                       const uint64_t* __restrict__ bucket, uint64_t* __restrict__ keys, uint32_t* __restrict__ point_list,
                       uint32_t* __restrict__ counters, SlabSource src, uint32_t* host_flags) {
    __shared__ uint64_t s[kSmallCap];
    if (counters[kAbortFlag] != 0u) return;
    const uint32_t tile = blockIdx.x;
    const uint2 r = ranges[tile];
    const int n = (int)(r.y - r.x), tid = threadIdx.x;
    if (n == 0 || n > kSmallCap) return;
    if (tid == 0 && cursor[tile] != r.y) flag_binning_mismatch(counters, host_flags);
    const int np = pow2_at_least(n);
    const EmitSorted emit{keys + r.x, point_list + r.x, (uint64_t)tile << 32, r.x, src};
    if (np <= 512)
        bitonic_sort_tile<2, 256>(bucket + r.x, n, np, s, tid, emit);
    else if (np == 1024)
        bitonic_sort_tile<4, 256>(bucket + r.x, n, np, s, tid, emit);
    else
        bitonic_sort_tile<8, 256>(bucket + r.x, n, np, s, tid, emit);
}

// merge of two sorted runs a[0,na) and b[0,nb) (unique keys): output element range [o0,o1) by merge path
// (plain pointers on purpose: the runs were written by this CTA in the previous pass, no read-only cache path)
__device__ __forceinline__ void merge_range(const uint64_t* a, int na, const uint64_t* b, int nb, uint64_t* out, int o0,
                                            int o1) {
    // co-rank: i = number of elements taken from a among the first o0 outputs
    int lo = max(0, o0 - nb), hi = min(o0, na);
    while (lo < hi) {
        const int i = (lo + hi) >> 1;  // candidate: i from a, o0-i from b
        if (a[i] < b[o0 - i - 1]) lo = i + 1; else hi = i;
    }
    int i = lo, j = o0 - lo;
    for (int o = o0; o < o1; ++o) {
        const bool take_a = (j >= nb) || (i < na && a[i] < b[j]);
        out[o] = take_a ? a[i++] : b[j++];
    }
}

// persistent CTAs over the list of long tiles: up to kLargeCap entries in shared memory, beyond that
// chunk sort + global merge passes (bucket <-> scratch)
__global__ void __launch_bounds__(kLargeThreads)
tile_sort_large_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ cursor,
                       const uint32_t* __restrict__ large_tiles, uint64_t* bucket, uint64_t* scratch, uint64_t* keys,
                       uint32_t* point_list,
                       uint32_t* __restrict__ counters, SlabSource src, uint32_t* host_flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
    const int tid = threadIdx.x;
    if (counters[kAbortFlag] != 0u) return;
    const uint32_t n_large = counters[3];
    for (uint32_t w = blockIdx.x; w < n_large; w += gridDim.x) {
        const uint32_t tile = large_tiles[w];
        const uint2 r = ranges[tile];
        const int n = (int)(r.y - r.x);
        if (tid == 0 && cursor[tile] != r.y) flag_binning_mismatch(counters, host_flags);
        if (n <= kLargeCap) {
            const int np = pow2_at_least(n);
            const EmitSorted emit{keys + r.x, point_list + r.x, (uint64_t)tile << 32, r.x, src};
            if (np <= 4096)
                bitonic_sort_tile<4, kLargeThreads>(bucket + r.x, n, np, s, tid, emit);
            else if (np == 8192)
                bitonic_sort_tile<8, kLargeThreads>(bucket + r.x, n, np, s, tid, emit);
            else
                bitonic_sort_tile<16, kLargeThreads>(bucket + r.x, n, np, s, tid, emit);
            __syncthreads();
            continue;
        }
        // chunk sort in place
        uint64_t* a = bucket + r.x;
        uint64_t* b = scratch + r.x;
        for (int c0 = 0; c0 < n; c0 += kLargeCap) {
            const int cn = min(kLargeCap, n - c0);
            bitonic_sort_tile<16, kLargeThreads>(a + c0, cn, kLargeCap, s, tid, EmitRaw{a + c0});
            __syncthreads();
        }
        // pairwise merge passes, every thread produces a contiguous slice of each merged pair
        for (int run = kLargeCap; run < n; run <<= 1) {
            __threadfence_block();
            __syncthreads();
            for (int p0 = 0; p0 < n; p0 += 2 * run) {
                const int na = min(run, n - p0), nb = max(0, min(run, n - p0 - run));
                const int total = na + nb;
                const int per = (total + kLargeThreads - 1) / kLargeThreads;
                const int o0 = min(total, tid * per), o1 = min(total, o0 + per);
                if (o0 < o1) merge_range(a + p0, na, a + p0 + run, nb, b + p0, o0, o1);
            }
            uint64_t* t = a;
            a = b;
            b = t;
        }
        __threadfence_block();
        __syncthreads();
        const EmitSorted emit{keys + r.x, point_list + r.x, (uint64_t)tile << 32, r.x, src};
        for (int i = tid; i < n; i += kLargeThreads) emit(i, a[i]);
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_duplicate(int P, const Frame& f, const Settings& s, const GeometryState& g, const int* radii,
                             const ImageState& img, const BinningState& b, size_t cap, cudaStream_t stream) {
    const int blocks = (P + 255) / 256;
    const bool tbc = s.tile_based_culling;
#define STP_DUP(TBC_, ORDER_) \
    duplicate_kernel<TBC_, ORDER_><<<blocks, 256, 0, stream>>>(P, f, g, radii, img.tile_cursor, b.bucket, (uint32_t)cap)
    switch (s.sort_order) {
        case 0:
        case 1:
            if (tbc) STP_DUP(true, 0); else STP_DUP(false, 0);
            break;
        case 2:
            if (tbc) STP_DUP(true, 2); else STP_DUP(false, 2);
            break;
        default:
            if (tbc) STP_DUP(true, 3); else STP_DUP(false, 3);
            break;
    }
#undef STP_DUP
    return cudaGetLastError();
}

cudaError_t launch_tile_scan(const Frame& f, const GeometryState& g, const ImageState& img, uint32_t capacity,
                             cudaStream_t stream) {
    const int tiles = f.grid_x * f.grid_y;
    tile_scan_kernel<<<(tiles + kScanThreads - 1) / kScanThreads, kScanThreads, 0, stream>>>(tiles, img.tile_count, img.ranges, img.tile_cursor,
                                                    img.large_tiles, g.counters, capacity);
    return cudaGetLastError();
}

int sort_kernel_launches() { return 2; }

cudaError_t launch_tile_sort(const Frame& f, const GeometryState& g, const ImageState& img, const BinningState& b,
                             const float* colors, uint32_t* host_flags, cudaStream_t stream) {
    // per call, not cached in a process-wide static: both the attribute and the SM count belong to the CURRENT device
    // (a process may drive several GPUs); the two driver calls cost about a microsecond
    int dev = 0, sm_count = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(tile_sort_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(kLargeCap * sizeof(uint64_t)));
    const SlabSource src{b.slab, b.slab_rgb, g.means2D, g.conic_opacity, g.cov3D_inv, colors};
    tile_sort_small_kernel<<<f.grid_x * f.grid_y, 256, 0, stream>>>(img.ranges, img.tile_cursor, b.bucket, b.keys,
                                                                    b.point_list, g.counters, src, host_flags);
    tile_sort_large_kernel<<<sm_count, kLargeThreads, kLargeCap * sizeof(uint64_t), stream>>>(
        img.ranges, img.tile_cursor, img.large_tiles, b.bucket, b.scratch, b.keys, b.point_list, g.counters, src, host_flags);
    return cudaGetLastError();
}

}  // namespace stp
