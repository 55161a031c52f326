// stp_kernels.cuh -- kernel argument packs and host-side launchers (one per pipeline stage).
#pragma once
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
    const float2* means2D;
    const float4* conic_opacity;
    const float4* cov3D_inv;
    const float* colors;  // [P,3] (SH-evaluated rgb or colors_precomp)
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    uint2* blend_rec;      // HIER blend log (nullptr = do not record)
    uint32_t* tile_flags;
    int rec_cap;
};

struct RenderBwdArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    const float2* means2D;
    const float4* conic_opacity;
    const float4* cov3D_inv;
    const float* colors;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* pixel_colors;
    const float* dL_dpix;
    float* dL_dmean2D;   // [P,3]
    float* dL_dconic;    // [P,4]
    float* dL_dopacity;  // [P]
    float* dL_dcolor;    // [P,3]
    const uint2* blend_rec;  // HIER blend log written by the forward pass (nullptr = re-sort everything)
    const uint32_t* tile_flags;
    int rec_cap;
};

struct PreprocessBwdArgs {
    int P, D, M;
    const float* means3D;
    const int* radii;
    const float* shs;
    const uint8_t* clamped;
    const float* opacities;
    const float* scales;
    const float* rotations;
    float scale_modifier;
    const float* cov3D;  // precomputed or geometry-state cov3D
    bool proper_ewa_scaling;
    const float* dL_dmean2D;
    const float* dL_dconic;
    float* dL_dopacity;
    float* dL_dmean3D;
    float* dL_dcolor;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
};

cudaError_t launch_preprocess(const PreprocessArgs& a, const Frame& f, const GeometryState& g, uint32_t* tile_count,
                              bool tbc, cudaStream_t stream);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t stream);

// binning.cu
int sort_kernel_launches();  // own kernels per tile sort
cudaError_t launch_tile_scan(const Frame& f, const GeometryState& g, const ImageState& img, cudaStream_t stream);
cudaError_t launch_duplicate(int P, const Frame& f, const Settings& s, const GeometryState& g, const int* radii,
                             const ImageState& img, const BinningState& b, size_t cap, cudaStream_t stream);
cudaError_t launch_tile_sort(const Frame& f, const GeometryState& g, const ImageState& img, const BinningState& b,
                             cudaStream_t stream);

// render_global.cu
cudaError_t launch_render_global_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_global_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream);

// render_hier.cu
cudaError_t launch_render_hier_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_hier_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream);

// render_ppx.cu
cudaError_t launch_render_kbuffer_fwd(const Frame& f, const Settings& s, const RenderArgs& a, cudaStream_t stream);
cudaError_t launch_render_kbuffer_bwd(const Frame& f, const Settings& s, const RenderBwdArgs& a, cudaStream_t stream);
cudaError_t launch_render_full_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream);

// preprocess_bwd.cu
cudaError_t launch_preprocess_bwd(const PreprocessBwdArgs& a, const Frame& f, cudaStream_t stream);

}  // namespace stp
