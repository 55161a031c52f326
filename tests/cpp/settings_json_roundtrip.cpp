// SplattingSettings <-> JSON through the converters of include/cuda_rasterizer/rasterizer.h (declared when nlohmann/json is
// included first): same schema as the reference (rasterizer.h:137-182), every key mandatory on input.
#include <nlohmann/json.hpp>
#include <rasterizer.h>
#include <cassert>
#include <iostream>
int main() {
    CudaRasterizer::SplattingSettings s{};
    s.sort_settings.sort_mode = CudaRasterizer::HIERARCHICAL;
    s.sort_settings.sort_order = CudaRasterizer::PER_TILE_DEPTH_MAXPOS;
    s.sort_settings.queue_sizes.per_pixel = 8;
    s.culling_settings.tile_based_culling = true;
    s.load_balancing = true;
    s.proper_ewa_scaling = false;
    nlohmann::json j = s;
    std::cout << j.dump() << std::endl;
    CudaRasterizer::SplattingSettings t = j.get<CudaRasterizer::SplattingSettings>();
    assert(t.sort_settings.sort_mode == s.sort_settings.sort_mode && t.sort_settings.queue_sizes.per_pixel == 8);
    assert(t.culling_settings.tile_based_culling && t.load_balancing && !t.proper_ewa_scaling);
    j.erase("load_balancing");
    bool threw = false;
    try { t = j.get<CudaRasterizer::SplattingSettings>(); } catch (const std::exception&) { threw = true; }
    assert(threw);
    return 0;
}
