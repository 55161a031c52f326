// rasterizer_shim.cu -- CudaRasterizer::Rasterizer (include/cuda_rasterizer/rasterizer.h) on top of the C ABI.
//
// Replaces: CudaRasterizer::Rasterizer::{markVisible, forward, backward} as declared in the reference's
// cuda_rasterizer/rasterizer.h:184-258 (implemented there in rasterizer_impl.cu:161-526), including the viewer-facing
// extras that only exist at this level: DebugVisualizationData (all six visualisation types, the statistics callback,
// rasterizer_impl.cu:54-109) and the stage timer whose report lands in timings_text every 128 frames
// (rasterizer_impl.h:77-147, rasterizer_impl.cu:248-249,387-399).
// Everything runs on the legacy default stream, like the reference.  Errors are thrown as std::runtime_error.
#include <cuda_runtime.h>

#include <mutex>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "../../include/cuda_rasterizer/rasterizer.h"
#include "../../include/stp_rasterizer.h"

namespace {

char* call_std_function(void* user, size_t bytes) { return (*static_cast<std::function<char*(size_t)>*>(user))(bytes); }

// size of every binning arena handed out through forward(), by address: backward() receives the bare pointer (and R,
// which this library does not need), the C ABI wants the arena size
std::mutex g_sizes_mutex;
std::unordered_map<const void*, size_t> g_binning_sizes;

struct BinningRecorder {
    std::function<char*(size_t)>* fn;
    char* last = nullptr;
    size_t bytes = 0;
};
char* call_and_record(void* user, size_t bytes) {
    BinningRecorder* r = static_cast<BinningRecorder*>(user);
    r->last = (*r->fn)(bytes);
    r->bytes = bytes;
    return r->last;
}

StpSettings make_settings(const CudaRasterizer::SortSettings& so, const CudaRasterizer::CullingSettings& cu, bool load_balancing,
                          bool proper_ewa_scaling) {
    StpSettings s{};
    s.sort_mode = static_cast<int>(so.sort_mode);
    s.sort_order = static_cast<int>(so.sort_order);
    s.queue_tile_4x4 = so.queue_sizes.tile_4x4;
    s.queue_tile_2x2 = so.queue_sizes.tile_2x2;
    s.queue_per_pixel = so.queue_sizes.per_pixel;
    s.rect_bounding = cu.rect_bounding;
    s.tight_opacity_bounding = cu.tight_opacity_bounding;
    s.tile_based_culling = cu.tile_based_culling;
    s.hierarchical_4x4_culling = cu.hierarchical_4x4_culling;
    s.load_balancing = load_balancing;
    s.proper_ewa_scaling = proper_ewa_scaling;
    return s;
}

int debug_code(DebugVisualization v) {
    switch (v) {
        case DebugVisualization::SortErrorOpacity: return STP_DEBUG_SORT_ERROR_OPACITY;
        case DebugVisualization::SortErrorDistance: return STP_DEBUG_SORT_ERROR_DISTANCE;
        case DebugVisualization::GaussianCountPerTile: return STP_DEBUG_COUNT_PER_TILE;
        case DebugVisualization::GaussianCountPerPixel: return STP_DEBUG_COUNT_PER_PIXEL;
        case DebugVisualization::Depth: return STP_DEBUG_DEPTH;
        case DebugVisualization::Transmittance: return STP_DEBUG_TRANSMITTANCE;
        default: return 0;
    }
}

// dL_dconic[P,4] = (x, y, -, w) of the packed accumulator (plane A [P][4]: conic.x, conic.y, conic.w, opacity)
__global__ void unpack_conic_kernel(int P, const float* __restrict__ acc, float* __restrict__ dL_dconic) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 a = *reinterpret_cast<const float4*>(acc + 4 * (size_t)i);  // plane A of the accumulator
    *reinterpret_cast<float4*>(dL_dconic + 4 * (size_t)i) = make_float4(a.x, a.y, 0.f, a.z);
}

constexpr int kTimerInterval = 128;  // rasterizer_impl.h:80
thread_local int g_timed_frames = 0;

}  // namespace

namespace CudaRasterizer {

void Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    if (stp_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<uint8_t*>(present), nullptr) != STP_OK)
        throw std::runtime_error(stp_last_error());
}

int Rasterizer::forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer,
                        std::function<char*(size_t)> imageBuffer, const int P, int D, int M, const float* background,
                        const int width, int height, const SplattingSettings splatting_settings,
                        DebugVisualizationData& debugVisualization, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales, const float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* inv_viewprojmatrix, const float* cam_pos, const float tan_fovx, float tan_fovy,
                        const bool prefiltered, float* out_color, int* radii, bool debug) {
    StpSettings s = make_settings(splatting_settings.sort_settings, splatting_settings.culling_settings,
                                  splatting_settings.load_balancing, splatting_settings.proper_ewa_scaling);
    s.debug_visualization = debug_code(debugVisualization.type);
    if (s.debug_visualization != 0) {
        s.blend_record_cap = 256;  // the visualisations replay the blend log (debug_vis.cu)
        s.debug_normalize = debugVisualization.debug_normalize;
        s.debug_min = debugVisualization.minMax[0];
        s.debug_max = debugVisualization.minMax[1];
        s.debug_pixel_x = debugVisualization.debugPixel[0];
        s.debug_pixel_y = debugVisualization.debugPixel[1];
    }
    int* radii_arg = radii;
    int* own_radii = nullptr;
    if (radii == nullptr && P > 0) {  // optional in the reference (rasterizer_impl.cu:262-265)
        if (cudaMalloc(&own_radii, sizeof(int) * (size_t)P) != cudaSuccess) throw std::runtime_error("cudaMalloc(radii) failed");
        radii_arg = own_radii;
    }
    BinningRecorder rec{&binningBuffer};
    int num_rendered = 0;
    const int flags = (debug ? 1 : 0) | (debugVisualization.timing_enabled ? 2 : 0);
    const int rc = stp_forward(call_std_function, &geometryBuffer, call_and_record, &rec, call_std_function, &imageBuffer, P, D,
                               M, background, width, height, &s, nullptr, means3D, shs, colors_precomp, opacities, scales,
                               scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, cam_pos,
                               tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color, radii_arg, flags, nullptr, &num_rendered);
    if (own_radii != nullptr) {
        cudaStreamSynchronize(nullptr);
        cudaFree(own_radii);
    }
    if (rc != STP_OK) throw std::runtime_error(stp_last_error());
    if (rec.last != nullptr) {
        std::lock_guard<std::mutex> lock(g_sizes_mutex);
        if (g_binning_sizes.size() > 4096) g_binning_sizes.clear();
        g_binning_sizes[rec.last] = rec.bytes;
    }
    if (debugVisualization.timing_enabled && ++g_timed_frames >= kTimerInterval) {
        g_timed_frames = 0;
        float ms[16];
        const char* names[16];
        int counts[16];
        const int n = stp_timing_summary(ms, names, counts, 16);
        std::stringstream ss;
        ss << "Timings: \n";
        float total = 0.f;
        for (int i = 0; i < n; ++i) {
            ss << " - " << names[i] << ": " << ms[i] << "ms\n";
            total += ms[i];
        }
        ss << " - Total: " << total << "ms\n";
        debugVisualization.timings_text = ss.str();
    }
    if (s.debug_visualization != 0) {
        float st[5];
        stp_last_debug_stats(st);
        debugVisualization.dataCallback(debugVisualization, st[0], st[1], st[2], st[3], st[4]);
    }
    return num_rendered;
}

void Rasterizer::backward(const int P, int D, int M, int /*R*/, const float* background, const int width, int height,
                          const SortSettings sort_settings, const CullingSettings culling_settings,
                          const bool proper_ewa_scaling, const float* means3D, const float* shs, const float* opacities,
                          const float* colors_precomp, const float* scales, const float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* inv_viewprojmatrix, const float* cam_pos, const float tan_fovx, float tan_fovy,
                          const float* pixel_colors, const int* radii, char* geom_buffer, char* binning_buffer,
                          char* image_buffer, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                          float* dL_drot, bool debug) {
    if (P <= 0) return;
    const StpSettings s = make_settings(sort_settings, culling_settings, false, proper_ewa_scaling);
    size_t binning_bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_sizes_mutex);
        auto it = g_binning_sizes.find(binning_buffer);
        if (it == g_binning_sizes.end())
            throw std::runtime_error("CudaRasterizer::Rasterizer::backward: binning_buffer was not produced by forward()");
        binning_bytes = it->second;
    }
    float* accum = nullptr;  // the packed screen-space accumulator of the C ABI (the reference zero-fills nine arrays)
    if (cudaMalloc(&accum, sizeof(float) * 9 * (size_t)P) != cudaSuccess) throw std::runtime_error("cudaMalloc(grad_accum) failed");
    cudaMemsetAsync(accum, 0, sizeof(float) * 9 * (size_t)P, nullptr);
    const int rc = stp_backward(P, D, M, binning_bytes, background, width, height, &s, nullptr, means3D, shs, opacities,
                                colors_precomp, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                                inv_viewprojmatrix, cam_pos, tan_fovx, tan_fovy, pixel_colors, radii, geom_buffer, binning_buffer,
                                image_buffer, dL_dpix, dL_dmean2D, accum, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
                                dL_dscale, dL_drot, debug ? 1 : 0, nullptr);
    if (rc == STP_OK && dL_dconic != nullptr) unpack_conic_kernel<<<(P + 255) / 256, 256>>>(P, accum, dL_dconic);
    cudaStreamSynchronize(nullptr);
    cudaFree(accum);
    if (rc != STP_OK) throw std::runtime_error(stp_last_error());
}

}  // namespace CudaRasterizer
