/*
 * stp_rasterizer.h -- C ABI of the B200-native sorted-Gaussian rasterizer (libstp_rasterizer.so).
 *
 * This is the drop-in boundary for the hot path of r4dl/StopThePop-Rasterization.  Every entry
 * point replaces one function of the reference's native interface; the citation after "replaces:"
 * is the reference file:line it stands in for.  Plain pointers and sizes only (no torch types):
 * all data pointers are DEVICE pointers to contiguous float32/int32 arrays unless stated
 * otherwise; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * Conventions kept from the reference (rasterize_points.cu:43-139, rasterizer.h:184-258):
 *   - matrices are 16 floats holding the TRANSPOSED 4x4 (i.e. glm column-major) matrix
 *   - an absent optional input is a NULL pointer (shs | colors_precomp, scales+rotations | cov3D_precomp)
 *   - the three scratch arenas (geometry / binning / image) are owned by the caller and obtained
 *     through an allocation callback, exactly like the std::function<char*(size_t)> arguments of
 *     CudaRasterizer::Rasterizer::forward (rasterizer.h:195-198); they are opaque and only valid
 *     as inputs of stp_backward / the stp_view_* decoders of the same library build.
 *
 * Error behaviour: every function returns 0 on success and a negative STP_ERR_* code on failure;
 * stp_last_error() returns a thread-local human-readable message (the reference throws
 * std::runtime_error with the same wording for unsupported queue sizes / PPX_FULL backward,
 * forward.cu:455-480, backward.cu:733-736).
 */
#ifndef STP_RASTERIZER_H_INCLUDED
#define STP_RASTERIZER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STP_ABI_VERSION 7

/* replaces: enum SortMode / GlobalSortOrder, rasterizer.h:27-41 */
enum { STP_SORT_GLOBAL = 0, STP_SORT_PPX_FULL = 1, STP_SORT_PPX_KBUFFER = 2, STP_SORT_HIER = 3 };
enum { STP_ORDER_Z_DEPTH = 0, STP_ORDER_DISTANCE = 1, STP_ORDER_PTD_CENTER = 2, STP_ORDER_PTD_MAX = 3 };

enum {
    STP_OK = 0,
    STP_ERR_INVALID_ARGUMENT = -1,
    STP_ERR_UNSUPPORTED = -2, /* e.g. queue size not instantiated, PPX_FULL backward of the reference */
    STP_ERR_CUDA = -3,
    STP_ERR_ALLOC = -4
};

/* replaces: struct SplattingSettings (+SortSettings, CullingSettings, SortQueueSizes),
 * rasterizer.h:43-135 and its from_json, rasterizer.h:160-182 (every key mandatory there). */
typedef struct StpSettings {
    int32_t sort_mode;   /* STP_SORT_*  */
    int32_t sort_order;  /* STP_ORDER_* */
    int32_t queue_tile_4x4; /* parsed but unused by the reference kernels (tail is hard-coded 64) */
    int32_t queue_tile_2x2; /* HIER mid queue: 8 | 12 | 20 */
    int32_t queue_per_pixel; /* HIER head queue: 4 | 8 | 16 (+12 bwd);  KBUFFER window 1..24 */
    int32_t rect_bounding;
    int32_t tight_opacity_bounding;
    int32_t tile_based_culling;
    int32_t hierarchical_4x4_culling;
    int32_t load_balancing;   /* scheduling hint only: results never depend on it */
    int32_t proper_ewa_scaling;
    /* not part of the reference's settings: GLOBAL / HIER modes, blend records per pixel kept by the forward pass for the
     * backward pass (8 B each, in the image arena; stp_image_bytes).  0 = none: backward repeats the re-sort.
     * Must have the same value in stp_forward and the matching stp_backward. */
    int32_t blend_record_cap;
    /* replaces DebugVisualizationData (rasterizer_debug.h:11-56): 0 = disabled, else one of STP_DEBUG_*: out_color becomes
     * the colour-mapped (Turbo for Depth, Magma for the others), min/max-normalised visualisation.  STP_DEBUG_DEPTH is what
     * render_depth=True of the Python API selects (rasterize_points.cu:104-107).  Forward only; all types except
     * COUNT_PER_TILE and TRANSMITTANCE need blend_record_cap > 0. */
    int32_t debug_visualization;
    int32_t debug_normalize;  /* normalise with [debug_min, debug_max] instead of the frame's own range (minMax) */
    float debug_min, debug_max;
    int32_t debug_pixel_x, debug_pixel_y; /* debugPixel: its raw value is reported by stp_last_debug_stats */
} StpSettings;
#define STP_DEBUG_SORT_ERROR_OPACITY 1  /* DebugVisualization::SortErrorOpacity      */
#define STP_DEBUG_SORT_ERROR_DISTANCE 2 /* DebugVisualization::SortErrorDistance     */
#define STP_DEBUG_COUNT_PER_TILE 3      /* DebugVisualization::GaussianCountPerTile  */
#define STP_DEBUG_DEPTH 4               /* DebugVisualization::Depth                 */
#define STP_DEBUG_COUNT_PER_PIXEL 5     /* DebugVisualization::GaussianCountPerPixel */
#define STP_DEBUG_TRANSMITTANCE 6       /* DebugVisualization::Transmittance         */

/* replaces: std::function<char*(size_t)> geometryBuffer/binningBuffer/imageBuffer,
 * rasterizer.h:195-198 (resizeFunctional, rasterize_points.cu:33-41).  Must return a device
 * pointer to at least `bytes` bytes, aligned to 256 B, or NULL on failure. */
typedef char* (*stp_alloc_fn)(void* user, size_t bytes);

/* Optional tile-row band for multi-GPU tile sharding.  Rows are 16-pixel tile rows;
 * [row_begin,row_end) = [0,-1] or a NULL pointer means the whole image.  With a band, Gaussians
 * are binned only into tiles of the band, only pixels of the band are written to out_color and
 * backward only consumes dL_dpix of the band: concatenating the bands of all ranks reproduces the
 * single-GPU point_list / ranges / image bit for bit. */
typedef struct StpTileBand {
    int32_t row_begin;
    int32_t row_end;
} StpTileBand;

/* replaces: CudaRasterizer::Rasterizer::forward, rasterizer.h:195-220 (impl rasterizer_impl.cu:221-413)
 * as called by RasterizeGaussiansCUDA, rasterize_points.cu:43-139.
 *   P Gaussians, D active SH degree, M SH coefficients per Gaussian (0 if shs==NULL).
 *   out_color  [3,H,W] f32 (every pixel of the image -- of the band, with a tile band -- is written; with a band the
 *              caller zero-fills the rest, like torch::full at rasterize_points.cu:80)
 *   radii      [P] i32     (written for every Gaussian; 0 = culled)
 *   num_rendered_out  HOST int: number of (tile,Gaussian) instances R (one stream sync, like
 *                     the reference's cudaMemcpy at rasterizer_impl.cu:317)
 *   debug             bit 0: synchronise + check after every stage (the reference's debug=true); bit 1: record stage
 *                     timings (stp_timing_summary); bit 2 = STP_FORWARD_ASYNC: do NOT synchronise.
 * STP_FORWARD_ASYNC: num_rendered_out must then point to TWO ints of PINNED host memory that stay valid until the stream
 * has passed the call: [0] receives R, [1] the device error flags (bit 0: prefiltered violation), both by an
 * asynchronous copy.  The binning arena is sized from the largest R this device has seen (x1.5; stp_note_num_rendered /
 * stp_set_num_rendered_hint) -- a device without history falls back to the synchronous path.  If the frame does not fit
 * (R > stp_binning_capacity(size of the binning arena)), every kernel after the tile scan returns immediately, out_color
 * is all zeros and the three arenas are invalid: the caller must check R once the stream has passed the call (before the
 * backward pass at the latest) and repeat the frame -- diff_gaussian_rasterization/_C.py does that on first use of R.
 */
#define STP_FORWARD_ASYNC 4
int stp_forward(stp_alloc_fn geom_alloc, void* geom_user,
                stp_alloc_fn binning_alloc, void* binning_user,
                stp_alloc_fn image_alloc, void* image_user,
                int P, int D, int M,
                const float* background, int width, int height,
                const StpSettings* settings, const StpTileBand* band,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix,
                const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, int* radii, int debug, void* stream,
                int* num_rendered_out);

/* replaces: CudaRasterizer::Rasterizer::backward, rasterizer.h:222-257 (impl rasterizer_impl.cu:417-526)
 * as called by RasterizeGaussiansBackwardCUDA, rasterize_points.cu:141-232.
 * grad_accum (9 P floats) is the only buffer the caller must zero-fill: the render-backward kernels accumulate the
 * screen-space gradients there, packed for 128-bit vector reductions in three planes (A [P][4]: conic.x, conic.y, conic.w,
 * opacity; B [P][4]: mean2D.x, mean2D.y, color.r, color.g; C [P]: color.b; the reference zero-fills nine separate tensors,
 * rasterize_points.cu:178-186).  It is also what a tile-sharded run all-reduces (36 B per Gaussian).
 * grad_accum and dL_drot must be 16-byte aligned (128-bit reductions / stores); shs and dL_dsh take a faster path
 * when they are.  All dL_* arrays are pure outputs and may be uninitialised -- every row is written, zeros for culled
 * Gaussians:
 * dL_dmean2D [P,3], dL_dopacity [P,1], dL_dcolor [P,3], dL_dmean3D [P,3], dL_dcov3D [P,6], dL_dsh [P,M,3],
 * dL_dscale [P,3], dL_drot [P,4].
 * binning_bytes: size of the binning arena as the allocation callback of stp_forward was (last) asked for it.  The
 * reference passes num_rendered (R) here to re-derive the carve-up (rasterizer_impl.cu:449); this library carves the
 * arena for the CAPACITY it allocated (stp_binning_capacity(binning_bytes) >= R), so the backward pass never needs R
 * on the host.
 */
int stp_backward(int P, int D, int M, size_t binning_bytes,
                 const float* background, int width, int height,
                 const StpSettings* settings, const StpTileBand* band,
                 const float* means3D, const float* shs, const float* opacities,
                 const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix,
                 const float* cam_pos, float tan_fovx, float tan_fovy,
                 const float* pixel_colors, const int* radii,
                 char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix,
                 float* dL_dmean2D, float* grad_accum, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                 int debug, void* stream);

/* The two halves of stp_backward as separate calls (same argument list), for data-parallel callers that overlap the
 * gradient exchange with the computation: stp_backward_render runs the render-backward stage (fills grad_accum);
 * stp_backward_preprocess then produces the dL_* rows of Gaussians [first, first+count) (first % 256 == 0), so the
 * all-reduce of one row range can run while the next range is computed (diff_gaussian_rasterization/_C.py,
 * sync_group=...).  stp_backward == stp_backward_render + stp_backward_preprocess(0, P). */
int stp_backward_render(int P, int D, int M, size_t binning_bytes,
                 const float* background, int width, int height,
                 const StpSettings* settings, const StpTileBand* band,
                 const float* means3D, const float* shs, const float* opacities,
                 const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix,
                 const float* cam_pos, float tan_fovx, float tan_fovy,
                 const float* pixel_colors, const int* radii,
                 char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix,
                 float* dL_dmean2D, float* grad_accum, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                 int debug, void* stream);
int stp_backward_preprocess(int P, int D, int M, size_t binning_bytes,
                 const float* background, int width, int height,
                 const StpSettings* settings, const StpTileBand* band,
                 const float* means3D, const float* shs, const float* opacities,
                 const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix,
                 const float* cam_pos, float tan_fovx, float tan_fovy,
                 const float* pixel_colors, const int* radii,
                 char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix,
                 float* dL_dmean2D, float* grad_accum, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                 int debug, void* stream, int first, int count);

/* replaces: CudaRasterizer::Rasterizer::markVisible, rasterizer.h:188-193 (rasterizer_impl.cu:161-173);
 * present is a device array of P bytes (bool). projmatrix is accepted and unused, as in the reference. */
int stp_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* Decoders of the opaque arenas (the reference's GeometryState/BinningState/ImageState::fromChunk,
 * rasterizer_impl.cu:175-217).  They fill a table of device pointers INTO the given buffer; used by
 * the parity tests to compare point_list / ranges / final_T / n_contrib / per-Gaussian scratch with
 * the reference build, and by stp_backward itself. */
typedef struct StpGeometryView {
    float* depths;           /* [P]   f32  view-space z or camera distance            */
    uint8_t* clamped;        /* [3P]  u8   SH->RGB clamp flags                         */
    float* rects2D;          /* [2P]  f32  rect extents (x,y)                          */
    float* means2D;          /* [2P]  f32  pixel-space mean                            */
    float* cov3D;            /* [6P]  f32  world covariance, upper triangle            */
    float* cov3D_inv;        /* [12P] f32  3 x float4 per Gaussian, NULL unless depth-along-ray is needed */
    float* conic_opacity;    /* [4P]  f32  conic (x,y,z) + opacity                     */
    float* rgb;              /* [3P]  f32  SH-evaluated colour                         */
    uint32_t* tiles_touched; /* [P]   u32                                              */
} StpGeometryView;
typedef struct StpBinningView {
    uint32_t* point_list;        /* [capacity] u32 sorted Gaussian ids, the first R valid  */
    uint64_t* point_list_keys;   /* [capacity] u64 sorted (tile<<32 | depth bits) keys     */
} StpBinningView;
typedef struct StpImageView {
    float* final_T;       /* [W*H] f32 */
    uint32_t* n_contrib;  /* [W*H] u32 (GLOBAL / KBUFFER / PPX_FULL only, like the reference) */
    uint32_t* ranges;     /* [2*tiles] u32 (start,end) per tile */
} StpImageView;

/* high-water mark of R on the current device, which sizes the speculative / asynchronous binning arena:
 * note = raise it to at least R (asynchronous callers report the R they resolved), set = overwrite it (0 = forget) */
void stp_note_num_rendered(int R);
void stp_set_num_rendered_hint(int R);

size_t stp_geometry_bytes(int P, int requires_cov3D_inv);
/* arena size for `capacity` instances (rounded up to a multiple of 64) under the given settings (the depth-resorting
 * modes add a 64-byte slab record per instance), and its exact inverse */
size_t stp_binning_bytes(int capacity, const StpSettings* settings);
int stp_binning_capacity(size_t binning_bytes, const StpSettings* settings);
size_t stp_image_bytes(int width, int height, int blend_record_cap);
int stp_view_geometry(char* geom_buffer, int P, int requires_cov3D_inv, StpGeometryView* out);
int stp_view_binning(char* binning_buffer, int capacity, StpBinningView* out);
int stp_view_image(char* image_buffer, int width, int height, StpImageView* out);
/* SortSettings::requiresDepthAlongRay, rasterizer.h:66-71 */
int stp_requires_cov3D_inv(const StpSettings* settings);

/* stage timings of the last stp_forward/stp_backward on this thread when debug&2 was set
 * (replaces the viewer-only Timer, rasterizer_impl.h:77-147): ms[0..n) with names[0..n). */
int stp_last_timings(float* ms, const char** names, int max_n);
/* Mean stage time over every debug&2 call on this thread since the last summary/reset: mean_ms[0..n),
 * names[0..n), counts[0..n).  Events are recorded on the caller's stream and resolved here, so the
 * instrumented calls themselves never synchronise (bench.py uses this for the per-kernel roofline). */
int stp_timing_summary(float* mean_ms, const char** names, int* counts, int max_n);
void stp_timing_reset(void);
/* cumulative number of hand-written kernels this thread has launched through stp_forward/stp_backward
 * (there is no library kernel on the path; memsets are not counted) -- bench.py reports the per-step difference. */
long long stp_kernel_launches(void);
/* statistics of the raw (pre-colormap) values of the last debug visualisation rendered on this thread -- the arguments
 * of DebugVisualizationData::dataCallback (rasterizer_impl.cu:97): value at the debug pixel, min, max, mean, std */
void stp_last_debug_stats(float* out5);

const char* stp_last_error(void);
int stp_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STP_RASTERIZER_H_INCLUDED */
