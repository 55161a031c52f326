// stp_state.cuh -- arena layouts (HBM data layout of the fwd->bwd scratch) and kernel argument packs.
//
// The three arenas play the role of the reference's GeometryState / BinningState / ImageState
// (rasterizer_impl.h:29-67, carve-up rasterizer_impl.cu:175-217): opaque byte buffers owned by the
// caller, re-derived from the same pointer in backward.  Sub-arrays are carved sequentially, each
// aligned to 256 B so that float4 / 128-bit accesses and TMA bulk copies (16 B granularity) are
// always legal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace stp {

constexpr size_t kArenaAlign = 256;
constexpr int kAbortFlag = 8;  // GeometryState::counters slot: the frame's instances do not fit the binning arena
constexpr int kPreprocessThreads = 256;

template <typename T>
static inline void obtain(char*& chunk, T*& ptr, size_t count) {
    size_t off = (reinterpret_cast<uintptr_t>(chunk) + kArenaAlign - 1) & ~(kArenaAlign - 1);
    ptr = reinterpret_cast<T*>(off);
    chunk = reinterpret_cast<char*>(ptr + count);
}

// per-Gaussian scratch: 83 B (+48 B with inverse covariance).  The reference's point_offsets / scan temp do not
// exist here: instance slots are claimed per tile (binning.cu), not per Gaussian.
struct GeometryState {
    float* depths;
    uint8_t* clamped;
    float2* rects2D;
    float2* means2D;
    float* cov3D;
    float4* cov3D_inv;  // nullptr unless requiresDepthAlongRay
    float4* conic_opacity;
    float* rgb;
    uint32_t* tiles_touched;
    uint32_t* counters;  // [1] R (total instances), [2] error flags, [3] number of tiles on the large-tile sort list,
                         // [4] pixels whose blend log overflowed in PPX_FULL mode, [5..7] depth visualisation: overflow
                         // count, min and max of the accumulated depth (order-preserving integer images),
                         // [kAbortFlag] set by the tile scan of an asynchronous forward whose binning arena is too small

    static GeometryState from_chunk(char*& chunk, size_t P, bool inv) {
        GeometryState g;
        obtain(chunk, g.depths, P);
        obtain(chunk, g.clamped, P * 3);
        obtain(chunk, g.rects2D, P);
        obtain(chunk, g.means2D, P);
        obtain(chunk, g.cov3D, P * 6);
        g.cov3D_inv = nullptr;
        if (inv) obtain(chunk, g.cov3D_inv, P * 3);
        obtain(chunk, g.conic_opacity, P);
        obtain(chunk, g.rgb, P * 3);
        obtain(chunk, g.tiles_touched, P);
        obtain(chunk, g.counters, 64);
        return g;
    }
};

struct ImageState {
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;
    // tile-bucket binning (binning.cu): per-tile instance histogram filled by preprocess, the running
    // slot cursor of every tile's bucket, and the list of tiles too long for the small in-smem sorter
    uint32_t* tile_count;
    uint32_t* tile_cursor;
    uint32_t* large_tiles;
    // Blend log (GLOBAL and HIER training steps): the forward pass logs every blend of every pixel, (Gaussian id, alpha),
    // at blend_rec[tile][k][thread] and the number of blends in blend_count; the backward pass replays the log front to
    // back instead of sweeping the tile list again (GLOBAL) / repeating the hierarchical re-sort (HIER).  Pixels with
    // more than rec_cap blends set tile_flags and are handled by the list-driven backward kernels.
    uint32_t* blend_count;
    uint32_t* tile_flags;
    uint2* blend_rec;  // nullptr when rec_cap == 0
    int rec_cap;
    static ImageState from_chunk(char*& chunk, size_t N, size_t tiles, int rec_cap) {
        ImageState s;
        obtain(chunk, s.final_T, N);
        obtain(chunk, s.n_contrib, N);
        obtain(chunk, s.ranges, tiles);
        obtain(chunk, s.tile_count, tiles);
        obtain(chunk, s.tile_cursor, tiles);
        obtain(chunk, s.large_tiles, tiles);
        obtain(chunk, s.tile_flags, tiles);
        obtain(chunk, s.blend_count, N);
        s.blend_rec = nullptr;
        s.rec_cap = rec_cap;
        if (rec_cap > 0) obtain(chunk, s.blend_rec, tiles * 256 * (size_t)rec_cap);
        return s;
    }
};

// per-instance arena: 28 B/instance (+80 B with slabs).  point_list / keys are the sorted outputs (the reference's
// point_list / point_list_keys, rasterizer_impl.h:57-67); bucket holds the unsorted
// (depth bits << 32 | Gaussian index) records grouped by tile, scratch is the ping-pong buffer of the
// global merge passes that only tiles longer than the in-smem capacity need; slab holds one 64-byte record per
// instance in list order (stp_slab.cuh) for the render modes that evaluate depth along rays.
// `cap` is the instance CAPACITY the arena was carved for (>= num_rendered); the backward pass re-derives it from the
// arena size (binning_capacity), so it never needs num_rendered on the host.
constexpr size_t kBinningGranule = 64;  // capacities are multiples of this: every sub-array is a whole number of 256-byte lines
static inline size_t binning_round_cap(size_t n) { return (n + kBinningGranule - 1) / kBinningGranule * kBinningGranule; }
struct BinningState {
    uint32_t* point_list;
    uint64_t* keys;
    uint64_t* bucket;
    uint64_t* scratch;
    float4* slab;      // nullptr unless requested
    float4* slab_rgb;  // [cap] {r, g, b, Gaussian id (bits)} in list order: what a blend needs besides the geometric record
    static BinningState from_chunk(char*& chunk, size_t cap, bool with_slab) {
        BinningState b;
        cap = binning_round_cap(cap);
        obtain(chunk, b.point_list, cap);
        obtain(chunk, b.keys, cap);
        obtain(chunk, b.bucket, cap);
        obtain(chunk, b.scratch, cap);
        b.slab = nullptr;
        b.slab_rgb = nullptr;
        if (with_slab) {
            obtain(chunk, b.slab, cap * 4);
            obtain(chunk, b.slab_rgb, cap);
        }
        return b;
    }
    static size_t bytes_per_instance(bool with_slab) { return 4 + 8 + 8 + 8 + (with_slab ? 64 + 16 : 0); }
};

template <typename T, typename... A>
static inline size_t required(A... a) {
    char* p = nullptr;
    T::from_chunk(p, a...);
    return reinterpret_cast<size_t>(p) + kArenaAlign;
}

// camera / frame constants, passed by value (lives in the constant bank of every kernel)
struct Frame {
    const float* viewmatrix;
    const float* projmatrix;
    const float* inv_viewproj;
    const float* cam_pos;
    const float* background;
    int W, H;
    int grid_x, grid_y;
    int row0, row1;  // tile-row band [row0,row1); whole image = [0,grid_y)
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
};

struct Settings {
    int sort_mode, sort_order;
    int q_mid, q_head;
    bool rect_bounding, tight_opacity_bounding, tile_based_culling, hier_culling, proper_ewa_scaling;
    int rec_cap;  // blend records per pixel (0 = none)
    bool render_depth;  // a debug visualisation instead of the colour image (render_depth=True of the Python API: Depth)
    int debug_vis;      // STP_DEBUG_* (0 = none)
    bool debug_normalize;
    float debug_min, debug_max;
    int debug_px, debug_py;
    bool per_tile_depth() const { return sort_order == 2 || sort_order == 3; }
    bool requires_inv() const { return sort_mode != 0 || per_tile_depth(); }
    bool uses_slab() const { return sort_mode != 0; }  // HIER / PPX_FULL / PPX_KBUFFER render from per-tile slabs
};

}  // namespace stp
