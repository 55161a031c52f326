// render_global.cu -- GLOBAL sort mode: front-to-back alpha blending of each tile's sorted list
// and the matching back-to-front gradient sweep.
//
// Replaces: renderCUDA<3,false> forward (forward.cu:234-366) and renderCUDA<3> backward
// (backward.cu:437-595).
//
// One CTA per 16x16 tile, one thread per pixel (pixel = tile origin + (tid%16, tid/16), the mapping
// whose per-pixel arithmetic we must reproduce).  The tile's list is streamed through shared memory
// in slabs of 256 entries {xy, conic+opacity, rgb}.  All per-(pixel,Gaussian) arithmetic is the
// rounding-pinned sequence of stp_math.cuh, so the accept/reject decisions (power>0, alpha<1/255,
// T<1e-4) and therefore final_T / n_contrib are identical to the reference build.
//
// Backward: every lane of a warp visits the same Gaussian in the same iteration, so the nine
// per-Gaussian gradient terms are first reduced across the warp with shuffles and only then
// accumulated per CTA in shared memory; one global float atomic per (tile, Gaussian, term) remains
// instead of one per (pixel, Gaussian, term) in the reference (backward.cu:561,583-592).
#include "stp_kernels.cuh"

namespace stp {

namespace {

constexpr int kTile = 16;
constexpr int kBlock = kTile * kTile;

__global__ void __launch_bounds__(kBlock)
render_global_fwd_kernel(Frame f, RenderArgs a) {
    __shared__ float2 s_xy[kBlock];
    __shared__ float4 s_co[kBlock];
    __shared__ float s_rgb[3][kBlock];

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const uint32_t px = tile_x * kTile + (tid & 15), py = tile_y * kTile + (tid >> 4);
    const bool inside = px < (uint32_t)f.W && py < (uint32_t)f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = a.ranges[tile_y * f.grid_x + tile_x];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + kBlock - 1) / kBlock;

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last_contributor = 0;

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        if (__syncthreads_count(done) == kBlock) break;
        const uint32_t src = range.x + r * kBlock + tid;
        if (src < range.y) {
            const uint32_t id = a.point_list[src];
            s_xy[tid] = a.means2D[id];
            s_co[tid] = a.conic_opacity[id];
            s_rgb[0][tid] = a.colors[3 * id + 0];
            s_rgb[1][tid] = a.colors[3 * id + 1];
            s_rgb[2][tid] = a.colors[3 * id + 2];
        }
        __syncthreads();
        const int n = min(kBlock, todo);
        for (int j = 0; !done && j < n; ++j) {
            ++contributor;
            const float2 xy = s_xy[j];
            const float4 co = s_co[j];
            const float dx = fsub(xy.x, pxf), dy = fsub(xy.y, pyf);
            const float power = gaussian_power(dx, dy, co.x, co.y, co.z);
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, fmul(co.w, expf(power)));
            if (alpha < kAlphaThreshold) continue;
            const float test_T = fmul(T, fsub(1.0f, alpha));
            if (test_T < kTThreshold) {
                done = true;
                continue;
            }
            C0 = ffma(T, fmul(alpha, s_rgb[0][j]), C0);
            C1 = ffma(T, fmul(alpha, s_rgb[1][j]), C1);
            C2 = ffma(T, fmul(alpha, s_rgb[2][j]), C2);
            T = test_T;
            last_contributor = contributor;
        }
    }

    if (inside) {
        a.final_T[pix_id] = T;
        a.n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)f.W * f.H;
        a.out_color[pix_id] = ffma(T, f.background[0], C0);
        a.out_color[plane + pix_id] = ffma(T, f.background[1], C1);
        a.out_color[2 * plane + pix_id] = ffma(T, f.background[2], C2);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kBlock)
render_global_bwd_kernel(Frame f, RenderBwdArgs a) {
    __shared__ uint32_t s_id[kBlock];
    __shared__ float2 s_xy[kBlock];
    __shared__ float4 s_co[kBlock];
    __shared__ float s_rgb[3][kBlock];
    __shared__ float s_acc[9][kBlock];  // per-slab gradient accumulators

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y + f.row0;
    const uint32_t px = tile_x * kTile + (tid & 15), py = tile_y * kTile + (tid >> 4);
    const bool inside = px < (uint32_t)f.W && py < (uint32_t)f.H;
    const uint32_t pix_id = (uint32_t)f.W * py + px;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = a.ranges[tile_y * f.grid_x + tile_x];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + kBlock - 1) / kBlock;

    const float T_final = inside ? a.final_T[pix_id] : 0.f;
    float T = T_final;
    uint32_t contributor = (uint32_t)todo;
    const uint32_t last_contributor = inside ? a.n_contrib[pix_id] : 0u;

    const size_t plane = (size_t)f.W * f.H;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (inside) {
        g0 = a.dL_dpix[pix_id];
        g1 = a.dL_dpix[plane + pix_id];
        g2 = a.dL_dpix[2 * plane + pix_id];
    }
    const float bg_dot = f.background[0] * g0 + f.background[1] * g1 + f.background[2] * g2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;

    for (int r = 0; r < rounds; ++r, todo -= kBlock) {
        __syncthreads();
        const int progress = r * kBlock + tid;
        if (range.x + progress < range.y) {
            const uint32_t id = a.point_list[range.y - progress - 1];
            s_id[tid] = id;
            s_xy[tid] = a.means2D[id];
            s_co[tid] = a.conic_opacity[id];
            s_rgb[0][tid] = a.colors[3 * id + 0];
            s_rgb[1][tid] = a.colors[3 * id + 1];
            s_rgb[2][tid] = a.colors[3 * id + 2];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) s_acc[k][tid] = 0.f;
        __syncthreads();
        const int n = min(kBlock, todo);
        for (int j = 0; j < n; ++j) {
            --contributor;
            bool hit = inside && contributor < last_contributor;
            float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
            float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hit) {
                const float2 xy = s_xy[j];
                co = s_co[j];
                dx = fsub(xy.x, pxf);
                dy = fsub(xy.y, pyf);
                const float power = gaussian_power(dx, dy, co.x, co.y, co.z);
                hit = !(power > 0.0f);
                if (hit) {
                    G = expf(power);
                    alpha = fminf(0.99f, fmul(co.w, G));
                    hit = !(alpha < kAlphaThreshold);
                }
            }
            if (!__any_sync(0xffffffffu, hit)) continue;
            float v[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = 0.f;
            if (hit) {
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                const float c0 = s_rgb[0][j], c1 = s_rgb[1][j], c2 = s_rgb[2][j];
                acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                lc0 = c0;
                lc1 = c1;
                lc2 = c2;
                float dL_dalpha = (c0 - acc0) * g0 + (c1 - acc1) * g1 + (c2 - acc2) * g2;
                v[0] = dchannel_dcolor * g0;
                v[1] = dchannel_dcolor * g1;
                v[2] = dchannel_dcolor * g2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = co.w * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * co.x - gdy * co.y;
                const float dG_ddely = -gdy * co.z - gdx * co.y;
                v[3] = dL_dG * dG_ddelx * ddelx_dx;
                v[4] = dL_dG * dG_ddely * ddely_dy;
                v[5] = -0.5f * gdx * dx * dL_dG;
                v[6] = -0.5f * gdx * dy * dL_dG;
                v[7] = -0.5f * gdy * dy * dL_dG;
                v[8] = G * dL_dalpha;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = warp_sum(v[k]);
            if (lane < 9) {
                float mine = v[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) mine = (lane == k) ? v[k] : mine;
                atomicAdd(&s_acc[lane][j], mine);
            }
        }
        __syncthreads();
        if (tid < n) {
            const uint32_t id = s_id[tid];
            const float c0 = s_acc[0][tid], c1 = s_acc[1][tid], c2 = s_acc[2][tid];
            if (c0 != 0.f) atomicAdd(&a.dL_dcolor[3 * id + 0], c0);
            if (c1 != 0.f) atomicAdd(&a.dL_dcolor[3 * id + 1], c1);
            if (c2 != 0.f) atomicAdd(&a.dL_dcolor[3 * id + 2], c2);
            const float m0 = s_acc[3][tid], m1 = s_acc[4][tid];
            if (m0 != 0.f) atomicAdd(&a.dL_dmean2D[3 * id + 0], m0);
            if (m1 != 0.f) atomicAdd(&a.dL_dmean2D[3 * id + 1], m1);
            const float k0 = s_acc[5][tid], k1 = s_acc[6][tid], k2 = s_acc[7][tid];
            if (k0 != 0.f) atomicAdd(&a.dL_dconic[4 * id + 0], k0);
            if (k1 != 0.f) atomicAdd(&a.dL_dconic[4 * id + 1], k1);
            if (k2 != 0.f) atomicAdd(&a.dL_dconic[4 * id + 3], k2);
            const float o = s_acc[8][tid];
            if (o != 0.f) atomicAdd(&a.dL_dopacity[id], o);
        }
    }
}

}  // namespace

cudaError_t launch_render_global_fwd(const Frame& f, const RenderArgs& a, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    render_global_fwd_kernel<<<grid, kBlock, 0, stream>>>(f, a);
    return cudaGetLastError();
}

cudaError_t launch_render_global_bwd(const Frame& f, const RenderBwdArgs& a, cudaStream_t stream) {
    dim3 grid(f.grid_x, f.row1 - f.row0, 1);
    if (grid.y == 0) return cudaSuccess;
    render_global_bwd_kernel<<<grid, kBlock, 0, stream>>>(f, a);
    return cudaGetLastError();
}

}  // namespace stp
