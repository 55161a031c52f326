// api.cu -- the C ABI (include/stp_rasterizer.h): stage sequencing and arena carve-up.
//
// Replaces: CudaRasterizer::Rasterizer::{forward,backward,markVisible} (rasterizer_impl.cu:161-173,
// 221-413, 417-526) and the json -> SplattingSettings conversion (rasterizer.h:160-182).
// Everything is enqueued on the caller's stream; the only host<->device synchronisation is the
// read-back of num_rendered that sizes the binning arena (the reference blocks in the same place,
// rasterizer_impl.cu:317).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/stp_rasterizer.h"
#include "stp_kernels.cuh"

using namespace stp;

namespace {

thread_local std::string g_error;
thread_local float g_debug_stats[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // value at the debug pixel, min, max, mean, std
thread_local long long g_launches = 0;  // hand-written kernels launched by this thread (stp_kernel_launches)
thread_local std::vector<std::pair<const char*, float>> g_timings;

// Device-raised error flags that no read-back of the regular pipeline covers (the tile sort detecting that the
// preprocess histogram and the duplicate kernel's emission disagree): one page of mapped pinned host memory, written by
// the kernels with plain system-scope stores only when something is wrong, inspected by the host at the start of the
// next call (and at the end of the same call with debug&1).  flags[0]: binning mismatch.
uint32_t* host_error_flags() {
    static uint32_t* flags = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 256, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            std::memset(p, 0, 256);
            flags = static_cast<uint32_t*>(p);
        } else {
            (void)cudaGetLastError();
        }
    });
    return flags;
}

// high-water mark of num_rendered per device: sizes the speculative binning-arena request
constexpr int kMaxDevices = 64;
std::atomic<uint32_t> g_last_R[kMaxDevices];

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
// sticky device-raised flags: reported (once) by the first call that sees them
int report_device_flags() {
    uint32_t* flags = host_error_flags();
    if (flags == nullptr) return 0;
    volatile uint32_t* v = flags;
    if (v[0] != 0u) {
        v[0] = 0u;
        return fail(STP_ERR_CUDA, "tile binning: the per-tile instance histogram and the emitted instances disagree "
                                  "(flag raised by a tile-sort kernel of this or an earlier call); the rendered output of "
                                  "that call is invalid");
    }
    return 0;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_error = std::string("CUDA error in ") + where + ": " + cudaGetErrorString(e);
    return STP_ERR_CUDA;
}

#define STP_CUDA(call, where)                             \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, where); \
        if (debug & 1) {                                  \
            e__ = cudaStreamSynchronize(stream);          \
            if (e__ != cudaSuccess) return cuda_fail(e__, where); \
        }                                                 \
    } while (0)

bool convert_settings(const StpSettings* in, Settings& s, std::string& err, bool backward) {
    if (in == nullptr) {
        err = "settings must not be NULL";
        return false;
    }
    if (in->sort_mode < 0 || in->sort_mode > 3) {
        err = "invalid sort_mode";
        return false;
    }
    if (in->sort_order < 0 || in->sort_order > 3) {
        err = "invalid sort_order";
        return false;
    }
    s.sort_mode = in->sort_mode;
    s.sort_order = in->sort_order;
    s.q_mid = in->queue_tile_2x2;
    s.q_head = in->queue_per_pixel;
    s.rect_bounding = in->rect_bounding != 0;
    s.tight_opacity_bounding = in->tight_opacity_bounding != 0;
    s.tile_based_culling = in->tile_based_culling != 0;
    s.hier_culling = in->hierarchical_4x4_culling != 0;
    s.proper_ewa_scaling = in->proper_ewa_scaling != 0;
    s.debug_vis = backward ? 0 : in->debug_visualization;
    s.render_depth = s.debug_vis != 0;  // out_color receives a visualisation instead of the colour image
    if (s.debug_vis < 0 || s.debug_vis > STP_DEBUG_TRANSMITTANCE) {
        err = "invalid debug_visualization";
        return false;
    }
    s.debug_normalize = in->debug_normalize != 0;
    s.debug_min = in->debug_min;
    s.debug_max = in->debug_max;
    s.debug_px = in->debug_pixel_x;
    s.debug_py = in->debug_pixel_y;
    const bool vis_needs_log = s.debug_vis != 0 && s.debug_vis != STP_DEBUG_COUNT_PER_TILE && s.debug_vis != STP_DEBUG_TRANSMITTANCE;
    // the k-buffer forward only writes a log for the visualisations (its backward re-sorts)
    s.rec_cap = ((in->sort_mode != STP_SORT_PPX_KBUFFER || vis_needs_log) && in->blend_record_cap > 0)
                    ? in->blend_record_cap : 0;
    if (vis_needs_log && s.rec_cap == 0) {
        err = "render_depth / debug visualisations need the blend log (blend_record_cap > 0)";
        return false;
    }
    if (s.sort_mode == STP_SORT_HIER) {
        // instantiated queue sizes, forward.cu:445-480 / backward.cu:739-767
        if (s.q_mid != 8 && s.q_mid != 12 && s.q_mid != 20) {
            err = "Not supported mid queue size " + std::to_string(s.q_mid);
            return false;
        }
        const bool head_ok = s.q_head == 4 || s.q_head == 8 || s.q_head == 16 || (backward && s.q_head == 12);
        if (!head_ok) {
            err = "Not supported head queue size " + std::to_string(s.q_head);
            return false;
        }
    }
    return true;
}

Frame make_frame(const float* background, int W, int H, const StpTileBand* band, const float* viewmatrix,
                 const float* projmatrix, const float* inv_viewproj, const float* cam_pos, float tan_fovx,
                 float tan_fovy) {
    Frame f;
    f.viewmatrix = viewmatrix;
    f.projmatrix = projmatrix;
    f.inv_viewproj = inv_viewproj;
    f.cam_pos = cam_pos;
    f.background = background;
    f.W = W;
    f.H = H;
    f.grid_x = (W + 15) / 16;
    f.grid_y = (H + 15) / 16;
    f.row0 = 0;
    f.row1 = f.grid_y;
    if (band != nullptr && band->row_end >= 0) {
        f.row0 = band->row_begin < 0 ? 0 : (band->row_begin > f.grid_y ? f.grid_y : band->row_begin);
        f.row1 = band->row_end > f.grid_y ? f.grid_y : band->row_end;
        if (f.row1 < f.row0) f.row1 = f.row0;
    }
    f.tan_fovx = tan_fovx;
    f.tan_fovy = tan_fovy;
    // rasterizer_impl.cu:251-252
    f.focal_y = H / (2.0f * tan_fovy);
    f.focal_x = W / (2.0f * tan_fovx);
    return f;
}

// Stage timer (replaces the viewer-only Timer, rasterizer_impl.h:77-147).  Armed by debug&2.  Events are
// recorded on the caller's stream and resolved LAZILY (stp_last_timings / stp_timing_summary), so an
// instrumented call adds a few event records but no host<->device synchronisation to the timed region.
struct StageSample {
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
};
thread_local std::vector<StageSample> g_pending;
thread_local std::vector<cudaEvent_t> g_event_pool;  // resolved events are reused: cudaEventCreate costs microseconds
constexpr size_t kMaxPending = 8192;

struct StageTimer {
    bool on;
    cudaStream_t stream;
    StageSample cur;
    StageTimer(bool on_, cudaStream_t s) : on(on_ && g_pending.size() < kMaxPending), stream(s) { mark(nullptr); }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        if (!g_event_pool.empty()) {
            e = g_event_pool.back();
            g_event_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        cudaEventRecord(e, stream);
        cur.ev.push_back(e);
        cur.names.push_back(name);
    }
    void finish() {
        if (!on) return;
        g_pending.push_back(std::move(cur));
        on = false;
    }
};

void resolve_sample(StageSample& smp, std::vector<std::pair<const char*, float>>& out) {
    out.clear();
    if (smp.ev.empty()) return;
    cudaEventSynchronize(smp.ev.back());
    for (size_t i = 1; i < smp.ev.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, smp.ev[i - 1], smp.ev[i]);
        out.push_back({smp.names[i], ms});
    }
}
void destroy_pending() {
    for (auto& smp : g_pending)
        for (auto e : smp.ev) g_event_pool.push_back(e);
    g_pending.clear();
}

}  // namespace

extern "C" {

const char* stp_last_error(void) { return g_error.c_str(); }
int stp_abi_version(void) { return STP_ABI_VERSION; }

int stp_last_timings(float* ms, const char** names, int max_n) {
    if (!g_pending.empty()) resolve_sample(g_pending.back(), g_timings);
    int n = 0;
    for (auto& t : g_timings) {
        if (n >= max_n) break;
        ms[n] = t.second;
        names[n] = t.first;
        ++n;
    }
    return n;
}

int stp_timing_summary(float* mean_ms, const char** names, int* counts, int max_n) {
    std::vector<std::pair<const char*, float>> one;
    std::vector<const char*> nm;
    std::vector<double> sum;
    std::vector<int> cnt;
    for (auto& smp : g_pending) {
        resolve_sample(smp, one);
        for (auto& t : one) {
            size_t k = 0;
            while (k < nm.size() && std::strcmp(nm[k], t.first) != 0) ++k;
            if (k == nm.size()) {
                nm.push_back(t.first);
                sum.push_back(0.0);
                cnt.push_back(0);
            }
            sum[k] += t.second;
            cnt[k] += 1;
        }
    }
    destroy_pending();
    int n = 0;
    for (size_t k = 0; k < nm.size() && n < max_n; ++k, ++n) {
        mean_ms[n] = (float)(sum[k] / cnt[k]);
        names[n] = nm[k];
        counts[n] = cnt[k];
    }
    return n;
}

void stp_timing_reset(void) { destroy_pending(); }
long long stp_kernel_launches(void) { return g_launches; }
void stp_last_debug_stats(float* out5) {
    if (out5 != nullptr) std::memcpy(out5, g_debug_stats, sizeof(g_debug_stats));
}

int stp_requires_cov3D_inv(const StpSettings* s) {
    if (s == nullptr) return 0;
    return s->sort_mode != STP_SORT_GLOBAL || s->sort_order == STP_ORDER_PTD_CENTER || s->sort_order == STP_ORDER_PTD_MAX;
}

void stp_set_num_rendered_hint(int R) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    g_last_R[dev].store(R > 0 ? (uint32_t)R : 0u, std::memory_order_relaxed);
}

void stp_note_num_rendered(int R) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || R <= 0) return;
    uint32_t cur = g_last_R[dev].load(std::memory_order_relaxed);
    while ((uint32_t)R > cur && !g_last_R[dev].compare_exchange_weak(cur, (uint32_t)R, std::memory_order_relaxed)) {
    }
}

size_t stp_geometry_bytes(int P, int inv) { return required<GeometryState>((size_t)P, inv != 0); }
size_t stp_binning_bytes(int capacity, const StpSettings* settings) {
    const bool slab = settings != nullptr && settings->sort_mode != STP_SORT_GLOBAL;
    return required<BinningState>((size_t)(capacity > 0 ? capacity : 0), slab);
}
// inverse of stp_binning_bytes: capacities are multiples of 64 instances, so every sub-array is a whole number of 256-byte
// lines and the arena size is an exact linear function of the capacity
int stp_binning_capacity(size_t bytes, const StpSettings* settings) {
    const bool slab = settings != nullptr && settings->sort_mode != STP_SORT_GLOBAL;
    const size_t fixed = required<BinningState>((size_t)0, slab);
    if (bytes <= fixed) return 0;
    return (int)((bytes - fixed) / (BinningState::bytes_per_instance(slab) * kBinningGranule) * kBinningGranule);
}
size_t stp_image_bytes(int W, int H, int rec_cap) {
    return required<ImageState>((size_t)W * H, (size_t)((W + 15) / 16) * ((H + 15) / 16), rec_cap > 0 ? rec_cap : 0);
}

int stp_view_geometry(char* buf, int P, int inv, StpGeometryView* out) {
    if (buf == nullptr || out == nullptr) return fail(STP_ERR_INVALID_ARGUMENT, "null argument");
    char* p = buf;
    GeometryState g = GeometryState::from_chunk(p, (size_t)P, inv != 0);
    out->depths = g.depths;
    out->clamped = g.clamped;
    out->rects2D = reinterpret_cast<float*>(g.rects2D);
    out->means2D = reinterpret_cast<float*>(g.means2D);
    out->cov3D = g.cov3D;
    out->cov3D_inv = reinterpret_cast<float*>(g.cov3D_inv);
    out->conic_opacity = reinterpret_cast<float*>(g.conic_opacity);
    out->rgb = g.rgb;
    out->tiles_touched = g.tiles_touched;
    return STP_OK;
}
int stp_view_binning(char* buf, int capacity, StpBinningView* out) {
    if (buf == nullptr || out == nullptr) return fail(STP_ERR_INVALID_ARGUMENT, "null argument");
    char* p = buf;
    BinningState b = BinningState::from_chunk(p, (size_t)capacity, false);  // the index buffers come first
    out->point_list = b.point_list;
    out->point_list_keys = b.keys;
    return STP_OK;
}
int stp_view_image(char* buf, int W, int H, StpImageView* out) {
    if (buf == nullptr || out == nullptr) return fail(STP_ERR_INVALID_ARGUMENT, "null argument");
    char* p = buf;
    ImageState s = ImageState::from_chunk(p, (size_t)W * H, (size_t)((W + 15) / 16) * ((H + 15) / 16), 0);
    out->final_T = s.final_T;
    out->n_contrib = s.n_contrib;
    out->ranges = reinterpret_cast<uint32_t*>(s.ranges);
    return STP_OK;
}

int stp_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* /*projmatrix*/,
                     uint8_t* present, void* stream_) {
    if (P <= 0) return STP_OK;
    cudaError_t e = launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_);
    if (e != cudaSuccess) return cuda_fail(e, "mark_visible");
    return STP_OK;
}

int stp_forward(stp_alloc_fn geom_alloc, void* geom_user, stp_alloc_fn binning_alloc, void* binning_user,
                stp_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width,
                int height, const StpSettings* settings, const StpTileBand* band, const float* means3D, const float* shs,
                const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                const float* inv_viewprojmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, int* radii, int debug, void* stream_, int* num_rendered_out) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_rendered_out) *num_rendered_out = 0;
    Settings s;
    std::string err;
    if (!convert_settings(settings, s, err, false)) return fail(STP_ERR_UNSUPPORTED, err);
    if (P <= 0) return STP_OK;  // rasterize_points.cu:93
    if (int rc = report_device_flags()) return rc;
    if (!geom_alloc || !binning_alloc || !image_alloc) return fail(STP_ERR_INVALID_ARGUMENT, "allocator callbacks required");
    if (shs == nullptr && colors_precomp == nullptr)
        return fail(STP_ERR_INVALID_ARGUMENT, "For non-RGB, provide precomputed Gaussian colors!");
    if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
        return fail(STP_ERR_INVALID_ARGUMENT, "provide scales+rotations or cov3D_precomp");
    if (s.requires_inv() && (scales == nullptr || rotations == nullptr))
        return fail(STP_ERR_INVALID_ARGUMENT, "depth-along-ray sort modes need scales and rotations (forward.cu:208-211)");

    Frame f = make_frame(background, width, height, band, viewmatrix, projmatrix, inv_viewprojmatrix, cam_pos, tan_fovx,
                         tan_fovy);
    const int tiles = f.grid_x * f.grid_y;
    const bool inv = s.requires_inv();
    if (s.rec_cap > 0 && (double)tiles * 256.0 * (double)s.rec_cap >= 4294967296.0) {
        // the log is indexed with 32 bits: very large images get the largest capacity that still fits; where the log is
        // only an optimisation (GLOBAL / HIER: the list-driven backward handles overflowing pixels, or everything) a
        // uselessly small one is dropped instead.  The backward pass re-derives the capacity from the arena size.
        const int fit = (int)(4294967295.0 / ((double)tiles * 256.0));
        const bool needed = s.sort_mode == STP_SORT_PPX_FULL || s.render_depth;
        s.rec_cap = (fit >= 16 || needed) ? fit : 0;
        if (needed && s.rec_cap <= 0)
            return fail(STP_ERR_INVALID_ARGUMENT, "image too large for the blend log (indexed with 32 bits)");
    }
    StageTimer timer((debug & 2) != 0, stream);

    char* gp = geom_alloc(geom_user, required<GeometryState>((size_t)P, inv));
    if (!gp) return fail(STP_ERR_ALLOC, "geometry arena allocation failed");
    GeometryState g = GeometryState::from_chunk(gp, (size_t)P, inv);
    char* ip = image_alloc(image_user, required<ImageState>((size_t)width * height, (size_t)tiles, s.rec_cap));
    if (!ip) return fail(STP_ERR_ALLOC, "image arena allocation failed");
    ImageState img = ImageState::from_chunk(ip, (size_t)width * height, (size_t)tiles, s.rec_cap);

    PreprocessArgs pa;
    pa.P = P; pa.D = D; pa.M = M;
    pa.means3D = means3D; pa.scales = scales; pa.rotations = rotations; pa.opacities = opacities;
    pa.shs = shs; pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
    pa.scale_modifier = scale_modifier;
    pa.sort_order = s.sort_order;
    pa.rect_bounding = s.rect_bounding;
    pa.tight_opacity_bounding = s.tight_opacity_bounding;
    pa.proper_ewa_scaling = s.proper_ewa_scaling;
    pa.prefiltered = prefiltered != 0;
    pa.radii = radii;
    STP_CUDA(launch_preprocess(pa, f, g, img.tile_count, s.tile_based_culling, stream), "preprocess");

    // The binning arena is sized by R, which is only known after the preprocess kernel + tile scan.  It is requested
    // for 1.5x the largest R this device has seen, while those kernels run (the allocation callback goes through the
    // caller's allocator, tens of microseconds), and carved for the CAPACITY that was allocated (stp_binning_capacity of
    // its size -- what the backward pass re-derives).
    //   synchronous (default): R is read back (one stream synchronisation, where the reference blocks too,
    //     rasterizer_impl.cu:317); a too small arena is simply requested again before anything uses it.
    //   asynchronous (debug & STP_FORWARD_ASYNC, only once the device has a history): nothing is read back here.  The
    //     tile scan compares R with the capacity on the device; if the arena is too small every later kernel of the frame
    //     returns at once, the image stays black, and the caller -- who finds R > capacity in num_rendered_out[0] once
    //     the copy below has landed -- runs the frame again (diff_gaussian_rasterization/_C.py does, on first use of R).
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<uint32_t>& last_R = g_last_R[dev >= 0 && dev < kMaxDevices ? dev : 0];
    const uint32_t seen = last_R.load(std::memory_order_relaxed);
    const bool slab = s.uses_slab();
    const bool async = (debug & STP_FORWARD_ASYNC) != 0 && seen != 0 && num_rendered_out != nullptr && !s.render_depth &&
                       (debug & 1) == 0;
    size_t cap = seen ? binning_round_cap((size_t)seen + seen / 2 + 4096) : 0;
    STP_CUDA(launch_tile_scan(f, g, img, async ? (uint32_t)cap : 0xFFFFFFFFu, stream), "tile scan");
    g_launches += 2;
    timer.mark("Preprocess");
    char* bp = cap ? binning_alloc(binning_user, required<BinningState>(cap, slab)) : nullptr;
    if (async && !bp) return fail(STP_ERR_ALLOC, "binning arena allocation failed");

    uint32_t R = 0;
    if (async) {
        // num_rendered_out[0] = R, [1] = error flags of the preprocess kernel; valid once the stream has passed this point
        cudaError_t e = cudaMemcpyAsync(num_rendered_out, g.counters + 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return cuda_fail(e, "num_rendered read-back");
        e = cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * (size_t)width * height, stream);  // an aborted frame is black
        if (e != cudaSuccess) return cuda_fail(e, "num_rendered read-back");
        R = (uint32_t)cap;  // launch decisions below only need "maybe non-empty"
    } else {
        uint32_t rf[2] = {0, 0};  // counters[1] = R, counters[2] = error flags raised by the preprocess kernel
        cudaError_t e = cudaMemcpyAsync(rf, g.counters + 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return cuda_fail(e, "num_rendered read-back");
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return cuda_fail(e, "num_rendered read-back");
        R = rf[0];
        if (rf[1] & 1u)  // the reference traps here (auxiliary.h:226-233)
            return fail(STP_ERR_INVALID_ARGUMENT, "Point is filtered although prefiltered is set. This shouldn't happen!");
        if (num_rendered_out) {
            num_rendered_out[0] = (int)R;
            if (debug & STP_FORWARD_ASYNC) num_rendered_out[1] = -2;  // asked for asynchronous, answered synchronously
        }
        if (R > seen) last_R.store(R, std::memory_order_relaxed);
        if (bp == nullptr || (size_t)R > cap) {
            cap = binning_round_cap((size_t)R);
            bp = binning_alloc(binning_user, required<BinningState>(cap, slab));
        }
        if (!bp) return fail(STP_ERR_ALLOC, "binning arena allocation failed");
    }
    BinningState b = BinningState::from_chunk(bp, cap, slab);

    if (R > 0) {
        STP_CUDA(launch_duplicate(P, f, s, g, radii, img, b, (size_t)R, stream), "duplicate");
        g_launches += 1;
    }
    timer.mark("Duplicate");
    if (R > 0) {
        STP_CUDA(launch_tile_sort(f, g, img, b, colors_precomp != nullptr ? colors_precomp : g.rgb, host_error_flags(), stream),
                 "tile sort");
        g_launches += sort_kernel_launches();
    }
    timer.mark("Sort");

    RenderArgs ra;
    ra.ranges = img.ranges;
    ra.point_list = b.point_list;
    ra.slab = b.slab;
    ra.slab_rgb = b.slab_rgb;
    ra.means2D = g.means2D;
    ra.conic_opacity = g.conic_opacity;
    ra.cov3D_inv = g.cov3D_inv;
    ra.colors = colors_precomp != nullptr ? colors_precomp : g.rgb;
    ra.final_T = img.final_T;
    ra.n_contrib = img.n_contrib;
    ra.out_color = out_color;
    ra.blend_rec = img.blend_rec;
    ra.blend_count = img.blend_count;
    ra.tile_flags = img.tile_flags;
    ra.log_overflow = g.counters + 4;
    ra.abort_flag = async ? g.counters + kAbortFlag : nullptr;
    ra.full_sort_ray = false;
    ra.log_is_position = s.sort_mode == STP_SORT_HIER || s.sort_mode == STP_SORT_PPX_FULL;
    ra.rec_cap = s.rec_cap;
    if (s.sort_mode == STP_SORT_GLOBAL) {
        STP_CUDA(launch_render_global_fwd(f, ra, stream), "render (GLOBAL)");
    } else if (s.sort_mode == STP_SORT_PPX_KBUFFER) {
        STP_CUDA(launch_render_kbuffer_fwd(f, s, ra, stream), "render (PPX_KBUFFER)");
    } else if (s.sort_mode == STP_SORT_PPX_FULL) {
        STP_CUDA(launch_render_full_fwd(f, ra, stream), "render (PPX_FULL)");
    } else {
        STP_CUDA(launch_render_hier_fwd(f, s, ra, stream), "render (HIER)");
    }
    g_launches += (s.sort_mode == STP_SORT_PPX_FULL) ? 2 : 1;
    timer.mark("Render");
    if (s.debug_vis != 0) {  // rasterizer_impl.cu:54-109, 402-413 (applyDebugVisualization)
        ra.full_sort_ray = s.sort_mode == STP_SORT_PPX_FULL;
        STP_CUDA(launch_debug_visualisation(f, ra, s, means3D, g.counters, stream), "debug visualisation");
        // raw values are in out_color now: statistics for the viewer's read-out, then the colormap
        uint32_t raw[9];  // counters[5..13]: overflow count, min, max (ordered bits), -, -, sum and sum of squares (doubles)
        struct { uint32_t overflowed, mn, mx; double s1, s2; } h;
        float at_pixel = 0.f;
        cudaError_t e = cudaMemcpyAsync(raw, g.counters + 5, sizeof(raw), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && s.debug_px > 0 && s.debug_px < width && s.debug_py > 0 && s.debug_py < height)
            e = cudaMemcpyAsync(&at_pixel, out_color + (size_t)width * s.debug_py + s.debug_px, sizeof(float),
                                cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return cuda_fail(e, "debug visualisation");
        h.overflowed = raw[0];
        h.mn = raw[1];
        h.mx = raw[2];
        std::memcpy(&h.s1, raw + 5, sizeof(double));
        std::memcpy(&h.s2, raw + 7, sizeof(double));
        if (h.overflowed != 0)
            return fail(STP_ERR_UNSUPPORTED, "render_depth: " + std::to_string(h.overflowed) +
                                                 " pixels blended more than blend_record_cap entries; raise "
                                                 "STP_BLEND_RECORD_CAP");
        const double n_pix = (double)width * (double)(std::min(f.row1 * 16, height) - std::min(f.row0 * 16, height));
        auto unorder = [](uint32_t u) {
            const uint32_t b = u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu);
            float v;
            std::memcpy(&v, &b, sizeof(v));
            return v;
        };
        const double mean = n_pix > 0 ? h.s1 / n_pix : 0.0, var = n_pix > 0 ? std::max(0.0, h.s2 / n_pix - mean * mean) : 0.0;
        g_debug_stats[0] = at_pixel;
        g_debug_stats[1] = unorder(h.mn);
        g_debug_stats[2] = unorder(h.mx);
        g_debug_stats[3] = (float)mean;
        g_debug_stats[4] = (float)std::sqrt(var);
        STP_CUDA(launch_debug_colormap(f, ra, s, g.counters, stream), "debug visualisation");
        g_launches += 2;
        timer.mark("DebugVisualisation");
    }
    timer.finish();
    if (debug & 1) {  // every stage has been synchronised: flags raised by this very call are visible
        if (int rc = report_device_flags()) return rc;
    }
    return STP_OK;
}

}  // extern "C"

namespace {
// which = 1: render backward only; 2: preprocess backward only (Gaussians [first, first+count)); 3: both
int backward_impl(int which, int first, int count, int P, int D, int M, size_t binning_bytes, const float* background, int width,
                  int height, const StpSettings* settings, const StpTileBand* band, const float* means3D, const float* shs,
                  const float* opacities, const float* colors_precomp, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const float* inv_viewprojmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                  const float* pixel_colors, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, float* dL_dmean2D, float* grad_accum, float* dL_dopacity, float* dL_dcolor,
                  float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug,
                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Settings s;
    std::string err;
    if (!convert_settings(settings, s, err, true)) return fail(STP_ERR_UNSUPPORTED, err);
    if (P <= 0) return STP_OK;  // rasterize_points.cu:191
    if (s.sort_mode == STP_SORT_PPX_FULL && s.rec_cap == 0)  // backward.cu:735; with the blend log of the forward pass
        return fail(STP_ERR_UNSUPPORTED, "Backward not supported for full per-pixel sort");  // it is supported (replay)
    if (!geom_buffer || !binning_buffer || !image_buffer) return fail(STP_ERR_INVALID_ARGUMENT, "null arena");
    if ((which & 2) && (first < 0 || (first & 255) != 0 || count < 0 || first + count > P))
        return fail(STP_ERR_INVALID_ARGUMENT, "preprocess-backward range must start at a multiple of 256 inside [0,P]");

    Frame f = make_frame(background, width, height, band, viewmatrix, projmatrix, inv_viewprojmatrix, cam_pos, tan_fovx,
                         tan_fovy);
    const int tiles = f.grid_x * f.grid_y;
    const bool inv = s.requires_inv();
    StageTimer timer((debug & 2) != 0, stream);
    char* gp = geom_buffer;
    GeometryState g = GeometryState::from_chunk(gp, (size_t)P, inv);
    char* bp = binning_buffer;
    const size_t R = (size_t)stp_binning_capacity(binning_bytes, settings);  // capacity the forward pass carved the arena for
    BinningState b = BinningState::from_chunk(bp, R, s.uses_slab());
    char* ip = image_buffer;
    ImageState img = ImageState::from_chunk(ip, (size_t)width * height, (size_t)tiles, s.rec_cap);

    if (which & 1) {
        RenderBwdArgs ra;
        ra.ranges = img.ranges;
        ra.point_list = b.point_list;
        ra.slab = b.slab;
        ra.slab_rgb = b.slab_rgb;
        ra.means2D = g.means2D;
        ra.conic_opacity = g.conic_opacity;
        ra.cov3D_inv = g.cov3D_inv;
        ra.colors = colors_precomp != nullptr ? colors_precomp : g.rgb;
        ra.final_T = img.final_T;
        ra.n_contrib = img.n_contrib;
        ra.pixel_colors = pixel_colors;
        ra.dL_dpix = dL_dpix;
        ra.grad_accum = grad_accum;
        ra.P = P;
        ra.blend_rec = img.blend_rec;
        ra.blend_count = img.blend_count;
        ra.tile_flags = img.tile_flags;
        ra.rec_cap = s.rec_cap;
        if (R > 0) {
            if (s.sort_mode == STP_SORT_GLOBAL) {
                STP_CUDA(launch_render_global_bwd(f, ra, stream), "render backward (GLOBAL)");
            } else if (s.sort_mode == STP_SORT_PPX_KBUFFER) {
                STP_CUDA(launch_render_kbuffer_bwd(f, s, ra, stream), "render backward (PPX_KBUFFER)");
            } else if (s.sort_mode == STP_SORT_PPX_FULL) {
                // no list-driven fallback exists for this mode: a pixel that blended more than the log holds is an error
                uint32_t overflowed = 0;
                cudaError_t e = cudaMemcpyAsync(&overflowed, g.counters + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
                if (e != cudaSuccess) return cuda_fail(e, "blend log overflow read-back");
                if (overflowed != 0)
                    return fail(STP_ERR_UNSUPPORTED, "PPX_FULL backward: " + std::to_string(overflowed) +
                                                         " pixels blended more than blend_record_cap entries; raise "
                                                         "STP_BLEND_RECORD_CAP");
                STP_CUDA(launch_render_full_bwd(f, ra, stream), "render backward (PPX_FULL)");
            } else {
                STP_CUDA(launch_render_hier_bwd(f, s, ra, stream), "render backward (HIER)");
            }
        }
        g_launches += (R > 0) * ((s.rec_cap > 0 && s.sort_mode != STP_SORT_PPX_FULL) ? 2 : 1);
        timer.mark("RenderBackward");
    }

    if (which & 2) {
        PreprocessBwdArgs pa;
        pa.P = P; pa.D = D; pa.M = M;
        pa.first = first; pa.P_end = first + count;
        pa.means3D = means3D; pa.radii = radii; pa.shs = shs; pa.opacities = opacities;
        pa.scales = scales; pa.rotations = rotations; pa.scale_modifier = scale_modifier;
        pa.cov3D = cov3D_precomp != nullptr ? cov3D_precomp : g.cov3D;
        pa.proper_ewa_scaling = s.proper_ewa_scaling;
        pa.dL_dmean2D = dL_dmean2D; pa.grad_accum = grad_accum; pa.dL_dopacity = dL_dopacity;
        pa.dL_dmean3D = dL_dmean3D; pa.dL_dcolor = dL_dcolor; pa.dL_dcov3D = dL_dcov3D; pa.dL_dsh = dL_dsh;
        pa.dL_dscale = dL_dscale; pa.dL_drot = dL_drot;
        STP_CUDA(launch_preprocess_bwd(pa, f, stream), "preprocess backward");
        g_launches += count > 0;
        timer.mark("PreprocessBackward");
    }
    timer.finish();
    return STP_OK;
}
}  // namespace

extern "C" {

#define STP_BWD_PARAMS                                                                                                       \
    int P, int D, int M, size_t binning_bytes, const float *background, int width, int height,                              \
        const StpSettings *settings,                                                                                        \
        const StpTileBand *band, const float *means3D, const float *shs, const float *opacities,                             \
        const float *colors_precomp, const float *scales, float scale_modifier, const float *rotations,                      \
        const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix, const float *inv_viewprojmatrix,       \
        const float *cam_pos, float tan_fovx, float tan_fovy, const float *pixel_colors, const int *radii,                   \
        char *geom_buffer, char *binning_buffer, char *image_buffer, const float *dL_dpix, float *dL_dmean2D,                \
        float *grad_accum, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D, float *dL_dcov3D, float *dL_dsh,         \
        float *dL_dscale, float *dL_drot, int debug, void *stream
#define STP_BWD_ARGS                                                                                                         \
    P, D, M, binning_bytes, background, width, height, settings, band, means3D, shs, opacities, colors_precomp, scales, scale_modifier, \
        rotations, cov3D_precomp, viewmatrix, projmatrix, inv_viewprojmatrix, cam_pos, tan_fovx, tan_fovy, pixel_colors,    \
        radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dmean2D, grad_accum, dL_dopacity, dL_dcolor,          \
        dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, debug, stream

int stp_backward(STP_BWD_PARAMS) { return backward_impl(3, 0, P, STP_BWD_ARGS); }
int stp_backward_render(STP_BWD_PARAMS) { return backward_impl(1, 0, 0, STP_BWD_ARGS); }
int stp_backward_preprocess(STP_BWD_PARAMS, int first, int count) { return backward_impl(2, first, count, STP_BWD_ARGS); }

}  // extern "C"
