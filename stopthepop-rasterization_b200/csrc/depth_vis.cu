// depth_vis.cu -- render_depth=True: the "Depth" debug visualisation, the only one reachable from the Python API
// (rasterize_points.cu:104-107).
//
// Replaces: the ENABLE_DEBUG_VIZ instantiations of the four render kernels for DebugVisualization::Depth
// (accumSortingErrorDepth / outputDebugVis, stopthepop_common.cuh:264-307), applyDebugVisualization's min/max
// reduction (rasterizer_impl.cu:54-109) and render_debug_CUDA<DEPTH> with the Turbo colormap (forward.cu:674-729,
// stopthepop_common.cuh:645-656).
//
// The reference compiles every render kernel a second time with the depth accumulation inside the blend loop.  Here
// the ordinary forward pass runs once with the blend log on and this file replays the log: per pixel
// depthAccum = sum depth_i * alpha_i * T_i over its blends in order, with the mode's own depth (GLOBAL: distance of
// the Gaussian's centre from the camera, forward.cu:337-339; per-pixel modes: depth along the pixel's ray, the sort key).
#include "stp_kernels.cuh"
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

// pixel_map: 0 GLOBAL strips, 1 HIER blocks/quads, 2 row-major (k-buffer, full sort); ray_depth: per-pixel modes
template <int PIXEL_MAP, bool RAY_DEPTH>
__global__ void __launch_bounds__(256)
depth_replay_kernel(Frame f, RenderArgs a, const float* __restrict__ means3D, uint32_t* __restrict__ counters) {
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
        uint32_t n = a.blend_count[pix_id];
        if (n > (uint32_t)a.rec_cap) {
            atomicAdd(counters + 5, 1u);  // reported by the host: the log is too short for this view
            n = (uint32_t)a.rec_cap;
        }
        const uint32_t tile_lin = (uint32_t)(tile_y * f.grid_x + tile_x);
        const uint2* __restrict__ rec = a.blend_rec + (size_t)tile_lin * a.rec_cap * 256 + tid;
        Vec3 ray{0.f, 0.f, 1.f};
        if constexpr (RAY_DEPTH) {
            const RayCam cam = make_raycam(f.inv_viewproj, f.cam_pos, f.W, f.H);
            ray = view_ray(cam, (float)px, (float)py);
        }
        for (uint32_t k = 0; k < n; ++k) {
            const uint2 r = __ldcs(rec + (size_t)k * 256);
            const int id = (int)r.x;
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
            acc += depth * alpha * T;
            T = fmul(T, fsub(1.0f, alpha));
        }
        const size_t plane = (size_t)f.W * f.H;
        a.out_color[pix_id] = acc;        // outputDebugVis, stopthepop_common.cuh:294-298
        a.out_color[plane + pix_id] = T;
    }
    // min / max of depthAccum over the image (cub::DeviceReduce::Min / Max in the reference)
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

// render_debug_CUDA<DEPTH = true>, forward.cu:674-714: alpha = clamp(depthAccum + T * max, min, max) / (max - min), Turbo
__global__ void depth_colormap_kernel(int N, const uint32_t* __restrict__ counters, float* __restrict__ out_color) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const float mn = from_ordered_bits(counters[6]), mx = from_ordered_bits(counters[7]);
    const float T = out_color[N + idx];
    const float x = fminf(fmaxf(out_color[idx] + T * mx, mn), mx) / (mx - mn);
    const float interp = fminf(fmaxf(x * 255.f, 0.f), 255.f);
    const int lo = x > 0 ? (int)interp : 0;
    const int hi = lo >= 255 ? 255 : lo + 1;
    const float diff = interp - (float)lo;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float v = kTurboLut[lo][ch] + (kTurboLut[hi][ch] - kTurboLut[lo][ch]) * diff;
        out_color[(size_t)ch * N + idx] = fminf(fmaxf(v, 0.f), 1.f);
    }
}

}  // namespace

cudaError_t launch_depth_visualisation(const Frame& f, const RenderArgs& a, int sort_mode, const float* means3D,
                                       uint32_t* counters, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    const uint32_t init[3] = {0u, 0xFFFFFFFFu, 0u};  // overflow count, min (ordered bits), max
    cudaError_t e = cudaMemcpyAsync(counters + 5, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    switch (sort_mode) {
        case 0: depth_replay_kernel<0, false><<<grid, 256, 0, stream>>>(f, a, means3D, counters); break;
        case 3: depth_replay_kernel<1, true><<<grid, 256, 0, stream>>>(f, a, means3D, counters); break;
        default: depth_replay_kernel<2, true><<<grid, 256, 0, stream>>>(f, a, means3D, counters); break;
    }
    const int N = f.W * f.H;
    depth_colormap_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, counters, a.out_color);
    return cudaGetLastError();
}

}  // namespace stp
