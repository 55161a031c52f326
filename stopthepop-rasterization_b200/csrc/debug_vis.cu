// debug_vis.cu -- the debug visualisations of the viewer (DebugVisualization, rasterizer_debug.h:11-20): Depth (what
// render_depth=True of the Python API selects, rasterize_points.cu:104-107), the two sort-error measures, Gaussian counts
// per tile / per pixel and transmittance.
//
// Replaces: the ENABLE_DEBUG_VIZ instantiations of the four render kernels (accumSortingErrorDepth / outputDebugVis,
// stopthepop_common.cuh:264-307), applyDebugVisualization's min/max reduction and statistics
// (rasterizer_impl.cu:54-109) and render_debug_CUDA with the Turbo / Magma colormaps (forward.cu:674-729,
// stopthepop_common.cuh:623-656).
//
// The reference compiles every render kernel a second time with the accumulation inside the blend loop.  Here the
// ordinary forward pass runs once with the blend log on and this file replays the log: per pixel, over its blends in
// order with the mode's own depth (GLOBAL: distance of the Gaussian's centre from the camera, forward.cu:337-339;
// per-pixel modes: depth along the pixel's ray, the sort key)
//   Depth              sum depth_i * alpha_i * T_i                      (+ T for the background term, Turbo)
//   SortErrorOpacity   sum of alpha_i over blends with depth_i <= max depth seen so far       (Magma)
//   SortErrorDistance  sum of |max depth so far - depth_i| over the same blends               (Magma)
//   Transmittance      1 - T_final;   GaussianCountPerTile: length of the tile's list;
//   GaussianCountPerPixel: Gaussians BLENDED into the pixel (the reference counts the list entries its loop visited,
//   including the ones it skipped -- a number that depends on its loop structure, not on the image).
// Only Depth can be reached through the reference's Python API.  Its output is compared with the reference build's
// (tests/test_gpu_parity.py, tests/golden/depth_vis.npz); the other five types are compared with the CPU oracle's
// restatement of the ENABLE_DEBUG_VIZ kernels, whose accumulators the Depth golden images pin
// (test_debug_visualisation_matches_cpu_oracle), and through their defining properties (tests/test_gpu_matrix.py).
#include "stp_kernels.cuh"
#include "stp_slab.cuh"
#include "stp_turbo_lut.cuh"

namespace stp {

namespace {

__device__ __forceinline__ uint32_t ordered_bits(float x) {  // monotone float -> uint32 for atomicMin / atomicMax
    const uint32_t b = __float_as_uint(x);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// Magma as a degree-6 polynomial per channel: M. Zucker's public-domain fit of matplotlib's table
// (shadertoy "Matplotlib colormaps", 2018), the same fit the reference evaluates (stopthepop_common.cuh:623-642)
__device__ __forceinline__ void magma(float x, float* rgb) {
    constexpr float c[7][3] = {{-0.002136485053939582f, -0.000749655052795221f, -0.005386127855323933f},
                               {0.2516605407371642f, 0.6775232436837668f, 2.494026599312351f},
                               {8.353717279216625f, -3.577719514958484f, 0.3144679030132573f},
                               {-27.66873308576866f, 14.26473078096533f, -13.64921318813922f},
                               {52.17613981234068f, -27.94360607168351f, 12.94416944238394f},
                               {-50.76852536473588f, 29.04658282127291f, 4.23415299384598f},
                               {18.65570506591883f, -11.48977351997711f, -5.601961508734096f}};
    x = fminf(fmaxf(x, 0.f), 1.f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float v = c[6][ch];
#pragma unroll
        for (int k = 5; k >= 0; --k) v = c[k][ch] + x * v;
        rgb[ch] = fminf(fmaxf(v, 0.f), 1.f);
    }
}

// pixel_map: 0 GLOBAL strips, 1 HIER blocks/quads, 2 row-major (k-buffer, full sort); ray_depth: per-pixel modes
template <int PIXEL_MAP, bool RAY_DEPTH>
__global__ void __launch_bounds__(256)
debug_replay_kernel(Frame f, RenderArgs a, int type, const float* __restrict__ means3D, uint32_t* __restrict__ counters,
                    double* __restrict__ moments) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    int px, py;
    if constexpr (PIXEL_MAP == 2) {
        px = tile_x * 16 + (tid & 15);
        py = tile_y * 16 + (tid >> 4);
    } else if constexpr (PIXEL_MAP == 1) {
        const int half = lane >> 4, hl = lane & 15;
        const int b = warp * 2 + half, q = hl >> 2, p = hl & 3;
        px = tile_x * 16 + (b & 3) * 4 + (q & 1) * 2 + (p & 1);
        py = tile_y * 16 + (b >> 2) * 4 + (q >> 1) * 2 + (p >> 1);
    } else {
        px = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
        py = tile_y * 16 + (warp >> 1) * 4 + (lane >> 3);
    }
    const bool inside = px < f.W && py < f.H;
    float acc = 0.f, T = 1.0f;
    if (inside) {
        const uint32_t pix_id = (uint32_t)f.W * py + px;
        const uint32_t tile_lin = (uint32_t)(tile_y * f.grid_x + tile_x);
        const size_t plane = (size_t)f.W * f.H;
        if (type == STP_DEBUG_COUNT_PER_TILE) {
            const uint2 r = a.ranges[tile_lin];
            acc = (float)(r.y - r.x);
        } else if (type == STP_DEBUG_TRANSMITTANCE) {
            acc = 1.0f - a.final_T[pix_id];
        } else {
            uint32_t n = a.blend_count[pix_id];
            if (n > (uint32_t)a.rec_cap) {
                atomicAdd(counters + 5, 1u);  // reported by the host: the log is too short for this view
                n = (uint32_t)a.rec_cap;
            }
            if (type == STP_DEBUG_COUNT_PER_PIXEL) {
                acc = (float)n;
            } else {
                const uint2* __restrict__ rec = a.blend_rec + (size_t)tile_lin * a.rec_cap * 256 + tid;
                Vec3 ray{0.f, 0.f, 1.f};
                if constexpr (RAY_DEPTH) {
                    const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
                    ray = PIXEL_MAP == 2 && a.full_sort_ray ? view_ray_xloop(cam, (float)px, (float)py) : view_ray(cam, (float)px, (float)py);
                }
                float cur = -3.402823466e+38f;
                const uint32_t first = a.ranges[tile_lin].x;
                for (uint32_t k = 0; k < n; ++k) {
                    const uint2 r = __ldcs(rec + (size_t)k * 256);
                    // the log stores the tile-local list position in the slab modes HIER / PPX_FULL, the id otherwise
                    const int id = a.log_is_position ? __float_as_int(__ldg(a.slab_rgb + first + r.x).w) : (int)r.x;
                    const float alpha = __uint_as_float(r.y);
                    float depth;
                    if constexpr (RAY_DEPTH) {
                        const float4 i0 = __ldg(a.cov3D_inv + 3 * id), i1 = __ldg(a.cov3D_inv + 3 * id + 1), i2 = __ldg(a.cov3D_inv + 3 * id + 2);
                        const float ic[6] = {i0.x, i0.y, i0.z, i1.x, i1.y, i1.z};
                        depth = depth_along_ray(ic, i2.x, i2.y, i2.z, ray);
                    } else {
                        const float dx = f.cam_pos[0] - means3D[3 * id], dy = f.cam_pos[1] - means3D[3 * id + 1],
                                    dz = f.cam_pos[2] - means3D[3 * id + 2];
                        depth = sqrtf(dx * dx + dy * dy + dz * dz);
                    }
                    // accumSortingErrorDepth, stopthepop_common.cuh:264-281
                    if (type == STP_DEBUG_DEPTH) {
                        acc += depth * alpha * T;
                    } else if (depth <= cur) {
                        acc += type == STP_DEBUG_SORT_ERROR_OPACITY ? alpha : fabsf(cur - depth);
                    }
                    cur = fmaxf(cur, depth);
                    T = fmul(T, fsub(1.0f, alpha));
                }
            }
        }
        a.out_color[pix_id] = acc;  // outputDebugVis, stopthepop_common.cuh:284-307
        if (type == STP_DEBUG_DEPTH) a.out_color[plane + pix_id] = T;
    }
    // mean / standard deviation of the raw values for the viewer's read-out (rasterizer_impl.cu:86-95 sums on the host)
    {
        double s1 = inside ? (double)acc : 0.0, s2 = inside ? (double)acc * (double)acc : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            atomicAdd(moments, s1);
            atomicAdd(moments + 1, s2);
        }
    }
    // min / max of the raw values over the image (cub::DeviceReduce::Min / Max in the reference)
    float lo = inside ? acc : 3.402823466e+38f, hi = inside ? acc : -3.402823466e+38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0 && lo <= hi) {
        atomicMin(counters + 6, ordered_bits(lo));
        atomicMax(counters + 7, ordered_bits(hi));
    }
}

// render_debug_CUDA, forward.cu:674-714: Depth: alpha = clamp(depthAccum + T * max, min, max) / (max - min), Turbo;
// the others: alpha = clamp(value, min, max) / (max - min), Magma.  min / max: of the frame, or the caller's (debug_normalize)
__global__ void debug_colormap_kernel(int N, int type, const uint32_t* __restrict__ counters, bool normalize, float nmin,
                                      float nmax, float* __restrict__ out_color) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    float mn = from_ordered_bits(counters[6]), mx = from_ordered_bits(counters[7]);
    if (normalize) {
        mn = nmin;
        mx = nmax;
    }
    float rgb[3];
    if (type == STP_DEBUG_DEPTH) {
        const float T = out_color[N + idx];
        const float x = fminf(fmaxf(out_color[idx] + T * mx, mn), mx) / (mx - mn);
        const float interp = fminf(fmaxf(x * 255.f, 0.f), 255.f);
        const int lo = x > 0 ? (int)interp : 0;
        const int hi = lo >= 255 ? 255 : lo + 1;
        const float diff = interp - (float)lo;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            rgb[ch] = fminf(fmaxf(kTurboLut[lo][ch] + (kTurboLut[hi][ch] - kTurboLut[lo][ch]) * diff, 0.f), 1.f);
    } else {
        magma(fminf(fmaxf(out_color[idx], mn), mx) / (mx - mn), rgb);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * N + idx] = rgb[ch];
}

}  // namespace

cudaError_t launch_debug_visualisation(const Frame& f, const RenderArgs& a, const Settings& s, const float* means3D,
                                       uint32_t* counters, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    const uint32_t init[3] = {0u, 0xFFFFFFFFu, 0u};  // overflow count, min (ordered bits), max
    cudaError_t e = cudaMemcpyAsync(counters + 5, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    double* moments = reinterpret_cast<double*>(counters + 10);  // counters[10..13]: sum, sum of squares (8-byte aligned)
    e = cudaMemsetAsync(moments, 0, 2 * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    switch (s.sort_mode) {
        case 0: debug_replay_kernel<0, false><<<grid, 256, 0, stream>>>(f, a, s.debug_vis, means3D, counters, moments); break;
        case 3: debug_replay_kernel<1, true><<<grid, 256, 0, stream>>>(f, a, s.debug_vis, means3D, counters, moments); break;
        default: debug_replay_kernel<2, true><<<grid, 256, 0, stream>>>(f, a, s.debug_vis, means3D, counters, moments); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_debug_colormap(const Frame& f, const RenderArgs& a, const Settings& s, const uint32_t* counters,
                                  cudaStream_t stream) {
    const int N = f.W * f.H;
    debug_colormap_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, s.debug_vis, counters, s.debug_normalize, s.debug_min,
                                                               s.debug_max, a.out_color);
    return cudaGetLastError();
}

}  // namespace stp
