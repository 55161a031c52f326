// binning.cu -- (tile | depth) key emission, device radix sort, per-tile ranges.
//
// Replaces: duplicateWithKeysCUDA (forward.cu:25-65), duplicateWithKeys_extended<TBC,LB,ORDER>
// (stopthepop_common.cuh:324-621), cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:344-352),
// identifyTileRanges (rasterizer_impl.cu:133-158).
//
// Key = (tile_id << 32) | float_bits(depth), value = Gaussian index, emitted row-major over the tile
// rectangle, Gaussians in index order (offsets from the fused scan in preprocess.cu).  Only key
// bits [0, 32+higher_msb(tiles)) take part in the sort; the sort is stable, so equal keys keep
// ascending Gaussian index -- the order the reference's stable CUB sort produces.
#include "stp_kernels.cuh"
#include "radix_sort.cuh"

namespace stp {

namespace {

constexpr int kSeqTiles = 8;
constexpr uint32_t kInvalidTile = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t make_key(uint32_t tile, float depth) {
    return ((uint64_t)tile << 32) | (uint64_t)__float_as_uint(depth);
}

struct DupGaussian {
    float2 xy;
    float4 co;       // conic + opacity (only read if TBC or PTD_MAX)
    float ic[6];     // inverse covariance (PTD only)
    float ux, uy, uz;
    float depth;     // global depth (Z / DISTANCE)
    float thr;
    int x0, y0, w, n;  // rect origin, width, tile count
    uint32_t off, off_end, idx;
};

// evaluates one tile of one Gaussian: returns whether a key is emitted and its depth.
template <bool TBC, int ORDER>
__device__ __forceinline__ bool eval_tile(const DupGaussian& gs, const RayCam& cam, int t, uint32_t grid_x,
                                          uint32_t& tile_id, float& depth) {
    const int tx = gs.x0 + t % gs.w, ty = gs.y0 + t / gs.w;
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

template <bool TBC, int ORDER>
__global__ void __launch_bounds__(256)
duplicate_kernel(int P, Frame f, GeometryState g, const int* __restrict__ radii, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ values) {
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
    gs.off = gs.off_end = 0;
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
        gs.off = (idx == 0) ? 0u : g.point_offsets[idx - 1];
        gs.off_end = g.point_offsets[idx];
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

    // sequential head of the rectangle
    uint32_t off = gs.off;
    for (int t = 0; t < min(gs.n, kSeqTiles); ++t) {
        uint32_t tile_id;
        float depth;
        if (eval_tile<TBC, ORDER>(gs, cam, t, (uint32_t)f.grid_x, tile_id, depth)) {
            if (off < gs.off_end) {
                keys[off] = make_key(tile_id, depth);
                values[off] = gs.idx;
            }
            ++off;
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
        o.off_end = __shfl_sync(0xffffffffu, gs.off_end, src);
        uint32_t o_off = __shfl_sync(0xffffffffu, off, src);
        for (int base = kSeqTiles; base < o.n; base += 32) {
            const int t = base + lane;
            uint32_t tile_id = 0;
            float depth = 0.f;
            const bool w = (t < o.n) && eval_tile<TBC, ORDER>(o, cam, t, (uint32_t)f.grid_x, tile_id, depth);
            const uint32_t m = __ballot_sync(0xffffffffu, w);
            const uint32_t pos = o_off + __popc(m & ((1u << lane) - 1u));
            if (w && pos < o.off_end) {
                keys[pos] = make_key(tile_id, depth);
                values[pos] = o.idx;
            }
            o_off += __popc(m);
        }
        if (lane == src) off = o_off;
    }

    // shortfall padding (stopthepop_common.cuh:504-508,615-619): cannot happen while preprocess and
    // duplicate share the same rounding-pinned tile test, kept for robustness.
    for (; off < gs.off_end; ++off) {
        keys[off] = make_key(kInvalidTile, 3.402823466e+38f);
        values[off] = 0xFFFFFFFFu;
    }
}

// identifyTileRanges, rasterizer_impl.cu:133-158 (ranges pre-zeroed)
__global__ void tile_ranges_kernel(uint32_t R, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = (uint32_t)(keys[i] >> 32);
    const bool valid = cur != kInvalidTile;
    if (i == 0) {
        if (valid) ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (cur != prev) {
            if (prev != kInvalidTile) ranges[prev].y = i;
            if (valid) ranges[cur].x = i;
        }
    }
    if (i == R - 1 && valid) ranges[cur].y = R;
}

}  // namespace

cudaError_t launch_duplicate(int P, const Frame& f, const Settings& s, const GeometryState& g, const int* radii,
                             uint64_t* keys, uint32_t* values, cudaStream_t stream) {
    const int blocks = (P + 255) / 256;
    const bool tbc = s.tile_based_culling;
#define STP_DUP(TBC_, ORDER_) duplicate_kernel<TBC_, ORDER_><<<blocks, 256, 0, stream>>>(P, f, g, radii, keys, values)
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

size_t sort_temp_bytes(size_t R) { return radix_sort_temp_bytes(R); }
int sort_kernel_launches(size_t R, int end_bit) { return radix_sort_kernel_launches(R, end_bit); }

cudaError_t launch_sort(BinningState& b, size_t R, int end_bit, cudaStream_t stream) {
    return radix_sort_pairs(b.sort_space, b.sort_bytes, b.keys_unsorted, b.keys, b.point_list_unsorted, b.point_list, R,
                            end_bit, stream);
}

cudaError_t launch_tile_ranges(size_t R, const uint64_t* keys, uint2* ranges, int tiles, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)tiles, stream);
    if (e != cudaSuccess) return e;
    if (R > 0) tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>((uint32_t)R, keys, ranges);
    return cudaGetLastError();
}

}  // namespace stp
