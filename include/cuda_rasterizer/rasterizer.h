/*
 * rasterizer.h -- the raw-pointer C++ interface CudaRasterizer::Rasterizer, source-compatible with
 * r4dl/StopThePop-Rasterization (cuda_rasterizer/rasterizer.h:24-258): same namespace, enum / struct / member
 * names, argument order and defaults, so that the SIBR viewer and other C++ callers link against this library
 * unchanged (CMake target CudaRasterizer, CMakeLists.txt).  Implemented in
 * stopthepop-rasterization_b200/csrc/rasterizer_shim.cu on top of the C ABI (include/stp_rasterizer.h).
 *
 * Differences a caller can observe:
 *  - kernels run on the legacy default stream like the reference's; errors are std::runtime_error with the
 *    reference's wording where it has one;
 *  - the JSON converters (rasterizer.h:137-182 of the reference) are only declared when nlohmann/json was
 *    included before this header (this repository vendors no third-party code);
 *  - backward(): R is accepted and ignored (the arena carve-up follows from the binning buffer itself);
 *    dL_dconic receives the accumulated conic gradients (x, y, -, w), as in the reference.
 */
#ifndef CUDA_RASTERIZER_H_INCLUDED
#define CUDA_RASTERIZER_H_INCLUDED

#include <functional>
#include <string>
#include <vector>

#include "stopthepop/rasterizer_debug.h"

namespace CudaRasterizer {

enum SortMode { GLOBAL = 0, PER_PIXEL_FULL = 1, PER_PIXEL_KBUFFER = 2, HIERARCHICAL = 3 };
enum GlobalSortOrder { VIEWSPACE_Z = 0, DISTANCE = 1, PER_TILE_DEPTH_CENTER = 2, PER_TILE_DEPTH_MAXPOS = 3 };

struct SortQueueSizes {
    int tile_4x4 = 64;
    int tile_2x2 = 8;
    int per_pixel = 4;
};

/* instantiated queue sizes (the UI of the viewer iterates over these) */
static const std::vector<int> per_pixel_queue_sizes{1, 2, 4, 8, 12, 16, 20, 24};
static const std::vector<int> twobytwo_tile_queue_sizes{8, 12, 20};
static const std::vector<int> per_pixel_queue_sizes_hier{4, 8, 16};

struct SortSettings {
    SortMode sort_mode = SortMode::GLOBAL;
    GlobalSortOrder sort_order = GlobalSortOrder::VIEWSPACE_Z;
    SortQueueSizes queue_sizes;
    bool requiresDepthAlongRay() const {
        return sort_mode != GLOBAL || sort_order == PER_TILE_DEPTH_CENTER || sort_order == PER_TILE_DEPTH_MAXPOS;
    }
    bool hasModifiableWindowSize() { return sort_mode == HIERARCHICAL || sort_mode == PER_PIXEL_KBUFFER; }
};

struct CullingSettings {
    bool rect_bounding = false;
    bool tight_opacity_bounding = false;
    bool tile_based_culling = false;
    bool hierarchical_4x4_culling = false;
};

inline std::string toString(SortMode m) {
    static const char* const n[] = {"GLOBAL", "FULL SORT", "KBUFFER", "HIERARCHICAL"};
    return (m >= GLOBAL && m <= HIERARCHICAL) ? n[m] : "";
}
inline bool isInvalidSortMode(int m) { return m < GLOBAL || m > HIERARCHICAL; }
inline std::string toString(GlobalSortOrder m) {
    static const char* const n[] = {"VIEWSPACE_Z", "DISTANCE", "PER_TILE_DEPTH_CENTER", "PER_TILE_DEPTH_MAXPOS"};
    return (m >= VIEWSPACE_Z && m <= PER_TILE_DEPTH_MAXPOS) ? n[m] : "";
}
inline bool isInvalidSortOrder(int m) { return m < VIEWSPACE_Z || m > PER_TILE_DEPTH_MAXPOS; }

struct SplattingSettings {
    SortSettings sort_settings;
    CullingSettings culling_settings;
    bool load_balancing;
    bool proper_ewa_scaling;
};

#ifdef NLOHMANN_JSON_VERSION_MAJOR
/* JSON schema of the settings files (same keys as the reference; every key mandatory on input).  One visitor lists the
 * leaves; to_json writes them, from_json reads them with .at(), i.e. a missing key throws. */
namespace detail {
template <class Settings, class Leaf>
inline void visit_settings(Settings& s, Leaf leaf) {
    leaf("sort_settings", nullptr, "sort_mode", s.sort_settings.sort_mode);
    leaf("sort_settings", nullptr, "sort_order", s.sort_settings.sort_order);
    leaf("sort_settings", "queue_sizes", "tile_4x4", s.sort_settings.queue_sizes.tile_4x4);
    leaf("sort_settings", "queue_sizes", "tile_2x2", s.sort_settings.queue_sizes.tile_2x2);
    leaf("sort_settings", "queue_sizes", "per_pixel", s.sort_settings.queue_sizes.per_pixel);
    leaf("culling_settings", nullptr, "rect_bounding", s.culling_settings.rect_bounding);
    leaf("culling_settings", nullptr, "tight_opacity_bounding", s.culling_settings.tight_opacity_bounding);
    leaf("culling_settings", nullptr, "tile_based_culling", s.culling_settings.tile_based_culling);
    leaf("culling_settings", nullptr, "hierarchical_4x4_culling", s.culling_settings.hierarchical_4x4_culling);
    leaf(nullptr, nullptr, "load_balancing", s.load_balancing);
    leaf(nullptr, nullptr, "proper_ewa_scaling", s.proper_ewa_scaling);
}
}  // namespace detail
inline void to_json(nlohmann::json& j, const SplattingSettings& s) {
    j = nlohmann::json::object();
    detail::visit_settings(s, [&j](const char* group, const char* sub, const char* key, const auto& value) {
        nlohmann::json* node = &j;
        if (group) node = &(*node)[group];
        if (sub) node = &(*node)[sub];
        (*node)[key] = value;
    });
}
inline void from_json(const nlohmann::json& j, SplattingSettings& s) {
    detail::visit_settings(s, [&j](const char* group, const char* sub, const char* key, auto& value) {
        const nlohmann::json* node = &j;
        if (group) node = &node->at(group);
        if (sub) node = &node->at(sub);
        node->at(key).get_to(value);
    });
}
#endif

class Rasterizer {
public:
    /* present[i] = Gaussian i passes the near-plane test of the frustum */
    static void markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present);

    /* returns num_rendered; the three arenas are requested through the callbacks (device memory, kept by the caller) */
    static int forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer,
                       std::function<char*(size_t)> imageBuffer,
                       const int P, int D, int M,
                       const float* background, const int width, int height,
                       const SplattingSettings splatting_settings, DebugVisualizationData& debugVisualization,
                       const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                       const float* scales, const float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix, const float* cam_pos,
                       const float tan_fovx, float tan_fovy, const bool prefiltered,
                       float* out_color, int* radii = nullptr, bool debug = false);

    /* gradients of everything forward() consumed; the geometry / binning / image arenas are the ones forward() filled */
    static void backward(const int P, int D, int M, int R,
                         const float* background, const int width, int height,
                         const SortSettings sort_settings, const CullingSettings culling_settings, const bool proper_ewa_scaling,
                         const float* means3D, const float* shs, const float* opacities, const float* colors_precomp,
                         const float* scales, const float scale_modifier, const float* rotations, const float* cov3D_precomp,
                         const float* viewmatrix, const float* projmatrix, const float* inv_viewprojmatrix, const float* cam_pos,
                         const float tan_fovx, float tan_fovy,
                         const float* pixel_colors, const int* radii,
                         char* geom_buffer, char* binning_buffer, char* image_buffer,
                         const float* dL_dpix,
                         float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                         float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                         bool debug);
};

}  // namespace CudaRasterizer

#endif
