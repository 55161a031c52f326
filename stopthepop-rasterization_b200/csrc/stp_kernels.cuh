// stp_kernels.cuh -- kernel argument packs and host-side launchers (one per pipeline stage).
#pragma once
#include "../../include/stp_rasterizer.h"
#include "stp_math.cuh"
#include "stp_state.cuh"

namespace stp {

struct PreprocessArgs {
    int P, D, M;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    float scale_modifier;
    int sort_order;
    bool rect_bounding, tight_opacity_bounding, proper_ewa_scaling, prefiltered;
    int* radii;
};

struct RenderArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    const float4* slab;   // per-tile slabs in list order (stp_slab.cuh), nullptr in GLOBAL mode
    const float4* slab_rgb;  // {r, g, b, id} per instance, list order
    const float2* means2D;
    const float4* conic_opacity;
    const float4* cov3D_inv;
    const float* colors;  // [P,3] (SH-evaluated rgb or colors_precomp)
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    uint2* blend_rec;      // blend log (nullptr = do not record): (entry, alpha) per blend, entry = tile-local list
                           // position in the slab modes HIER / PPX_FULL, Gaussian id in GLOBAL / PPX_KBUFFER
    uint32_t* blend_count;
    uint32_t* tile_flags;
    uint32_t* log_overflow;  // counter: pixels whose log overflowed in a mode without list-driven backward (PPX_FULL)
    const uint32_t* abort_flag;  // non-zero: the binning arena of this (asynchronous) frame was too small -- render nothing
    bool full_sort_ray;          // debug visualisation of PPX_FULL: depths on the ray as that mode's kernels round it
    bool log_is_position;        // the blend log holds tile-local list positions (HIER / PPX_FULL), not Gaussian ids
    int rec_cap;
};

struct RenderBwdArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    const float4* slab;
    const float4* slab_rgb;
    const float2* means2D;
    const float4* conic_opacity;
    const float4* cov3D_inv;
    const float* colors;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* pixel_colors;
    const float* dL_dpix;
    float* grad_accum;   // packed screen-space gradient accumulator, 9 P floats in three planes (zero-filled by the caller)
    int P;               // number of Gaussians (plane stride of the accumulator)
    const uint2* blend_rec;  // blend log written by the forward pass (nullptr = list-driven backward for everything)
    const uint32_t* blend_count;
    const uint32_t* tile_flags;
    int rec_cap;
};

struct PreprocessBwdArgs {
    int P, D, M;
    int first, P_end;  // Gaussian range [first, P_end) of this launch (first % 256 == 0)
    const float* means3D;
    const int* radii;
    const float* shs;
    const float* opacities;
    const float* scales;
    const float* rotations;
    float scale_modifier;
    const float* cov3D;  // precomputed or geometry-state cov3D
    bool proper_ewa_scaling;
    const float* grad_accum;  // 9 P floats in three planes, filled by the render-backward kernels
    float* dL_dmean2D;        // [P,3] out (z = 0)
    float* dL_dopacity;       // [P]   out
    float* dL_dmean3D;
    float* dL_dcolor;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
};

// Packed accumulator of the screen-space gradients: nine floats per Gaussian in three PLANES
//   A [P][4]: dL_dconic.x, dL_dconic.y, dL_dconic.w, dL_dopacity     B [P][4]: dL_dmean2D.x, dL_dmean2D.y, dL_dcolor.r, .g
//   C [P]   : dL_dcolor.b
// so that one blend costs two 128-bit vector reductions (REDG.E.ADD.F32x4, sm_90+) and one scalar reduction instead of
// the reference's nine scalar float atomics (backward.cu:561,583-592): the render-backward kernels are bound by the
// number of reduction requests the SM can issue, not by the bytes.  Planes instead of 48-byte rows: every float of the
// buffer is used, so a tile-sharded run all-reduces 36 B per Gaussian, not 48.
constexpr int kGradAccumFloats = 9;
#ifdef __CUDACC__
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void accumulate_grads(float* __restrict__ acc, int P, int id, float col0, float col1, float col2,
                                                 float m0, float m1, float k0, float k1, float k2, float op) {
    red_add_v4(acc + 4 * (size_t)id, k0, k1, k2, op);
    red_add_v4(acc + 4 * ((size_t)P + id), m0, m1, col0, col1);
    atomicAdd(acc + 8 * (size_t)P + id, col2);
}
#endif

#ifdef __CUDACC__
// SH degree 3 (48 floats per Gaussian): the 32 coefficient rows of a warp are one contiguous span of rows*12 float4.
// Twelve fully unrolled steps, lane L moving float4 number L + 32 t: every float4 lies inside one row (48 % 4 == 0),
// so there is no per-element bookkeeping, and the independent 128-bit loads of several steps are in flight together.
// Shared-memory rows have an odd stride (49 floats) so that "lane r reads coefficient k of row r" is conflict-free.
template <bool STORE>
__device__ __forceinline__ void move_sh_rows_48(float* __restrict__ gptr, int rows, uint32_t live, int lane,
                                                float* __restrict__ my_rows) {
    float4* __restrict__ g4 = reinterpret_cast<float4*>(gptr);
    const int n4 = rows * 12;
#pragma unroll
    for (int t = 0; t < 12; ++t) {
        const int i = lane + 32 * t;
        const int row = i / 12, col = (i - row * 12) * 4;
        if (i < n4 && ((live >> row) & 1u)) {
            float* const dst = my_rows + row * 49 + col;
            if constexpr (STORE) {
                g4[i] = make_float4(dst[0], dst[1], dst[2], dst[3]);
            } else {
                const float4 v = __ldg(g4 + i);
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            }
        }
    }
}
#endif

cudaError_t launch_preprocess(const PreprocessArgs& a, const Frame& f, const GeometryState& g, uint32_t* tile_count,
                              bool tbc, cudaStream_t stream);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t stream);

// binning.cu
int sort_kernel_launches();  // own kernels per tile sort
cudaError_t launch_tile_scan(const Frame& f, const GeometryState& g, const ImageState& img, uint32_t capacity,
                             cudaStream_t stream);
cudaError_t launch_duplicate(int P, const Frame& f, const Settings& s, const GeometryState& g, const int* radii,
                             const ImageState& img, const BinningState& b, size_t cap, cudaStream_t stream);
cudaError_t launch_tile_sort(const Frame& f, const GeometryState& g, const ImageState& img, const BinningState& b,
                             const float* colors, uint32_t* host_flags, cudaStream_t stream);

// render_global.cu
cudaError_t launch_render_global_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_global_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream);

// render_hier.cu
cudaError_t launch_render_hier_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_hier_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream);

// blend-log replay backward (render_hier.cu): pixel_map selects the thread -> pixel map of the kernel that wrote the
// log (0 GLOBAL strips, 1 HIER blocks/quads, 2 row-major = PPX_FULL); whole_tile_fallback = skip every pixel of a flagged tile (GLOBAL) instead of only the overflowed pixels (HIER)
cudaError_t launch_blend_replay_bwd(const Frame& f, const RenderBwdArgs& a, int pixel_map, bool whole_tile_fallback,
                                    cudaStream_t stream);

// render_ppx.cu
cudaError_t launch_render_kbuffer_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_kbuffer_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream);
cudaError_t launch_render_full_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_full_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream);

// debug_vis.cu: DebugVisualization types (render_depth=True = Depth) from the blend log of the forward pass: raw values
// + min / max / moments into out_color / counters, then the colormap
cudaError_t launch_debug_visualisation(const Frame& f, const RenderArgs& a, const Settings& s, const float* means3D,
                                       uint32_t* counters, cudaStream_t stream);
cudaError_t launch_debug_colormap(const Frame& f, const RenderArgs& a, const Settings& s, const uint32_t* counters,
                                  cudaStream_t stream);

// preprocess_bwd.cu
cudaError_t launch_preprocess_bwd(const PreprocessBwdArgs& a, const Frame& f, cudaStream_t stream);

}  // namespace stp
