// C++ consumer of CudaRasterizer::Rasterizer (include/cuda_rasterizer/rasterizer.h), written the way a viewer calls the
// reference: std::function arenas, raw device pointers, DebugVisualizationData.  Reads a scene dumped by the Python
// test, renders it (forward, optionally backward), writes image / radii / one gradient for comparison with the
// Python path.   usage: shim_smoke <scene.bin> <out.bin> <sort_mode> <debug_type 0..6>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include <rasterizer.h>

static std::vector<float> read_floats(FILE* f, size_t n) {
    std::vector<float> v(n);
    if (fread(v.data(), sizeof(float), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}
template <typename T>
static T* to_device(const std::vector<T>& h) {
    T* d = nullptr;
    cudaMalloc(&d, sizeof(T) * h.size());
    cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    return d;
}
struct Arena {
    char* ptr = nullptr;
    size_t cap = 0;
    std::function<char*(size_t)> fn() {
        return [this](size_t n) {
            if (n > cap) { cudaFree(ptr); cudaMalloc(&ptr, n); cap = n; }
            return ptr;
        };
    }
};

int main(int argc, char** argv) {
    if (argc < 5) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int hdr[4];  // P, W, H, M
    if (fread(hdr, sizeof(int), 4, f) != 4) return 2;
    const int P = hdr[0], W = hdr[1], H = hdr[2], M = hdr[3];
    float tan[2];
    if (fread(tan, sizeof(float), 2, f) != 2) return 2;
    auto means = read_floats(f, 3 * (size_t)P), scales = read_floats(f, 3 * (size_t)P), rots = read_floats(f, 4 * (size_t)P),
         opac = read_floats(f, P), shs = read_floats(f, 3 * (size_t)M * P), view = read_floats(f, 16), proj = read_floats(f, 16),
         inv = read_floats(f, 16), campos = read_floats(f, 3), bg = read_floats(f, 3), dL = read_floats(f, 3 * (size_t)W * H);
    fclose(f);
    float *d_means = to_device(means), *d_scales = to_device(scales), *d_rots = to_device(rots), *d_opac = to_device(opac),
          *d_shs = to_device(shs), *d_view = to_device(view), *d_proj = to_device(proj), *d_inv = to_device(inv),
          *d_cam = to_device(campos), *d_bg = to_device(bg), *d_dL = to_device(dL);
    float* d_out = nullptr;
    int* d_radii = nullptr;
    cudaMalloc(&d_out, sizeof(float) * 3 * W * H);
    cudaMalloc(&d_radii, sizeof(int) * P);

    CudaRasterizer::SplattingSettings st{};
    st.sort_settings.sort_mode = static_cast<CudaRasterizer::SortMode>(atoi(argv[3]));
    st.load_balancing = false;
    st.proper_ewa_scaling = false;
    DebugVisualizationData dbg;
    dbg.type = static_cast<DebugVisualization>(atoi(argv[4]));
    float stats[5] = {0, 0, 0, 0, 0};
    dbg.dataCallback = [&](const DebugVisualizationData&, float v, float mn, float mx, float avg, float sd) {
        stats[0] = v; stats[1] = mn; stats[2] = mx; stats[3] = avg; stats[4] = sd;
    };
    dbg.timing_enabled = true;
    Arena geom, binning, img;
    int R = 0;
    for (int it = 0; it < 130; ++it)  // > 128 frames: the timer report is due
        R = CudaRasterizer::Rasterizer::forward(geom.fn(), binning.fn(), img.fn(), P, 3, M, d_bg, W, H, st, dbg, d_means, d_shs,
                                                nullptr, d_opac, d_scales, 1.0f, d_rots, nullptr, d_view, d_proj, d_inv, d_cam,
                                                tan[0], tan[1], false, d_out, d_radii, false);
    std::vector<float> out(3 * (size_t)W * H), gmean(3 * (size_t)P, 0.f);
    std::vector<int> radii(P);
    cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(radii.data(), d_radii, sizeof(int) * P, cudaMemcpyDeviceToHost);
    if (dbg.type == DebugVisualization::Disabled && st.sort_settings.sort_mode != CudaRasterizer::PER_PIXEL_FULL) {
        float *g2 = nullptr, *gc = nullptr, *go = nullptr, *gcol = nullptr, *g3 = nullptr, *gcov = nullptr, *gsh = nullptr,
              *gs = nullptr, *gr = nullptr;
        cudaMalloc(&g2, 12 * (size_t)P); cudaMalloc(&gc, 16 * (size_t)P); cudaMalloc(&go, 4 * (size_t)P);
        cudaMalloc(&gcol, 12 * (size_t)P); cudaMalloc(&g3, 12 * (size_t)P); cudaMalloc(&gcov, 24 * (size_t)P);
        cudaMalloc(&gsh, 12 * (size_t)M * P); cudaMalloc(&gs, 12 * (size_t)P); cudaMalloc(&gr, 16 * (size_t)P);
        CudaRasterizer::Rasterizer::backward(P, 3, M, R, d_bg, W, H, st.sort_settings, st.culling_settings, false, d_means, d_shs,
                                             d_opac, nullptr, d_scales, 1.0f, d_rots, nullptr, d_view, d_proj, d_inv, d_cam,
                                             tan[0], tan[1], d_out, d_radii, geom.ptr, binning.ptr, img.ptr, d_dL, g2, gc, go,
                                             gcol, g3, gcov, gsh, gs, gr, false);
        cudaMemcpy(gmean.data(), g3, sizeof(float) * gmean.size(), cudaMemcpyDeviceToHost);
    }
    FILE* o = fopen(argv[2], "wb");
    fwrite(&R, sizeof(int), 1, o);
    fwrite(stats, sizeof(float), 5, o);
    const int has_timings = dbg.timings_text.find("Render") != std::string::npos;
    fwrite(&has_timings, sizeof(int), 1, o);
    fwrite(out.data(), sizeof(float), out.size(), o);
    fwrite(radii.data(), sizeof(int), radii.size(), o);
    fwrite(gmean.data(), sizeof(float), gmean.size(), o);
    fclose(o);
    printf("R=%d %s", R, dbg.timings_text.c_str());
    return 0;
}
