/*
 * rasterizer_debug.h -- debug-visualisation request / result block of the CudaRasterizer C++ interface.
 *
 * Source-compatible with the type names of r4dl/StopThePop-Rasterization
 * (cuda_rasterizer/stopthepop/rasterizer_debug.h:11-56) so that a viewer written against the reference
 * (SIBR: DebugVisualizationData is filled by the UI and passed to Rasterizer::forward) compiles unchanged;
 * implemented by stopthepop-rasterization_b200/csrc/debug_vis.cu through the C ABI (stp_rasterizer.h).
 */
#pragma once

#include <functional>
#include <string>

enum class DebugVisualization { SortErrorOpacity, SortErrorDistance, GaussianCountPerTile, GaussianCountPerPixel, Depth, Transmittance, Disabled };

inline std::string toString(DebugVisualization m) {
    static const char* const names[] = {"Sort Error: Opacity",      "Sort Error: Distance", "Gaussian Count Per Tile",
                                        "Gaussian Count Per Pixel", "Depth",                "Transmittance"};
    const int i = static_cast<int>(m);
    return (i >= 0 && i < 6) ? names[i] : "Disabled";
}

struct DebugVisualizationData {
    DebugVisualization type{DebugVisualization::Disabled};
    int debugPixel[2] = {};
    /* called by Rasterizer::forward with (this, value at debugPixel, min, max, mean, standard deviation of the frame) */
    std::function<void(const DebugVisualizationData&, float, float, float, float, float)> dataCallback{
        [](const DebugVisualizationData&, float, float, float, float, float) {}};
    float minMax[2] = {0.f, 10000.f};
    bool debug_normalize = false; /* normalise with minMax instead of the frame's own range */

    std::string timings_text = ""; /* filled every 128 frames while timing_enabled */
    bool timing_enabled = false;
};

namespace sortQualityDebug {
inline bool isSortError(DebugVisualization v) { return v == DebugVisualization::SortErrorDistance || v == DebugVisualization::SortErrorOpacity; }
inline bool isVisualized(DebugVisualization v) { return v != DebugVisualization::Disabled; }
inline bool isMagma(DebugVisualization v) { return isVisualized(v) && v != DebugVisualization::Depth; }
}  // namespace sortQualityDebug
