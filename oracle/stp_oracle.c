/*
 * stp_oracle.c -- plain-C, single-file CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it.  The product (stopthepop-rasterization_b200/) has no CPU path.
 *
 * Parity status: PINNED -- validated against tests/golden/*.npz, which are outputs of the unmodified
 * reference CUDA build (oracle/_ref) run on a B200 (tests/golden/make_golden.py; the debug-visualisation
 * accumulators against its render_depth images, tests/golden/make_golden_depth_vis.py).  The reference
 * itself ships no CPU implementation, tests or golden vectors.
 * Integer outputs (radii, point_list, ranges, n_contrib) are restated exactly; floats differ from
 * the GPU only through libm expf/logf vs. CUDA's (<= 2 ulp), which can flip a threshold decision
 * for isolated (pixel, Gaussian) pairs -- the tests bound the count of such flips.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Float expressions that feed integer decisions spell out the FMA contraction of the reference
 * build with fmaf() (see DESIGN.md "Arithmetic contract"); compile with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define ALPHA_THRESHOLD (1.0f / 255.0f) /* auxiliary.h:21 */
#define T_THRESHOLD 0.0001f             /* auxiliary.h:23 */

typedef struct {
    int sort_mode, sort_order, q_tile4, q_mid, q_head;
    int rect_bounding, tight_opacity_bounding, tile_based_culling, hier_culling, load_balancing, proper_ewa_scaling;
} OrcSettings;

typedef struct {
    int P, D, M, W, H;
    const float *means3D, *scales, *rotations, *opacities, *shs, *colors_precomp, *cov3D_precomp;
    float scale_modifier;
    const float *viewmatrix, *projmatrix, *inv_viewproj, *campos, *bg;
    float tan_fovx, tan_fovy;
} OrcInputs;

/* per-Gaussian state = GeometryState (rasterizer_impl.h:29-46) */
typedef struct {
    int P, R, tiles, W, H;
    int *radii;
    float *depths, *means2D, *rects2D, *conic_opacity, *rgb, *cov3D, *cov3D_inv;
    uint8_t *clamped;
    uint32_t *tiles_touched, *point_offsets;
    uint64_t *keys;
    uint32_t *point_list;
    uint32_t *ranges; /* [tiles][2] */
    float *out_color, *final_T;
    uint32_t *n_contrib;
} OrcState;

static float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

/* ---- preprocess pieces ------------------------------------------------------------------------ */
/* computeCov3D rotation part, forward_common.h:165-169 (quaternion NOT normalised, :158) */
static void quat_to_rot(const float* q, float R[3][3]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float xz = x * z, zz = z * z, yy = y * y, rx = r * x, rz = r * z;
    const float xz_p = fmaf(r, y, xz), xz_m = fmaf(-r, y, xz);
    const float yz_m = fmaf(y, z, -rx), yz_p = fmaf(y, z, rx);
    const float xy_m = fmaf(x, y, -rz), xy_p = fmaf(x, y, rz);
    const float yy_zz = yy + zz, xx_yy = fmaf(x, x, yy), xx_zz = fmaf(x, x, zz);
    R[0][0] = 1.0f - (yy_zz + yy_zz); R[0][1] = xy_m + xy_m;           R[0][2] = xz_p + xz_p;
    R[1][0] = xy_p + xy_p;           R[1][1] = 1.0f - (xx_zz + xx_zz); R[1][2] = yz_m + yz_m;
    R[2][0] = xz_m + xz_m;           R[2][1] = yz_p + yz_p;           R[2][2] = 1.0f - (xx_yy + xx_yy);
}
/* Sigma = (S R)^T (S R), forward_common.h:171-182; also Sigma^-1 with S^-1, stopthepop_common.cuh:13-41 */
static void gram(const float R[3][3], float s0, float s1, float s2, float* o) {
    float m[3][3];
    for (int c = 0; c < 3; ++c) { m[c][0] = s0 * R[c][0]; m[c][1] = s1 * R[c][1]; m[c][2] = s2 * R[c][2]; }
    o[0] = dot3c(m[0][0], m[0][0], m[0][1], m[0][1], m[0][2], m[0][2]);
    o[1] = dot3c(m[0][0], m[1][0], m[0][1], m[1][1], m[0][2], m[1][2]);
    o[2] = dot3c(m[0][0], m[2][0], m[0][1], m[2][1], m[0][2], m[2][2]);
    o[3] = dot3c(m[1][0], m[1][0], m[1][1], m[1][1], m[1][2], m[1][2]);
    o[4] = dot3c(m[1][0], m[2][0], m[1][1], m[2][1], m[1][2], m[2][2]);
    o[5] = dot3c(m[2][0], m[2][0], m[2][1], m[2][1], m[2][2], m[2][2]);
}
/* computeCov2D, forward_common.h:73-106 */
static void cov2d(const float* pv, float fx, float fy, float tfx, float tfy, const float* c, const float* vm, float* out3) {
    const float limx = 1.3f * tfx, limy = 1.3f * tfy;
    const float txtz = pv[0] / pv[2], tytz = pv[1] / pv[2];
    const float tx = fminf(limx, fmaxf(-limx, txtz)) * pv[2], ty = fminf(limy, fmaxf(-limy, tytz)) * pv[2];
    const float tz2 = pv[2] * pv[2];
    const float J00 = fx / pv[2], J11 = fy / pv[2], J02 = -(fx * tx) / tz2, J12 = -(fy * ty) / tz2;
    const float T00 = fmaf(vm[2], J02, vm[0] * J00), T01 = fmaf(vm[6], J02, vm[4] * J00), T02 = fmaf(vm[10], J02, vm[8] * J00);
    const float T10 = fmaf(vm[2], J12, vm[1] * J11), T11 = fmaf(vm[6], J12, vm[5] * J11), T12 = fmaf(vm[10], J12, vm[9] * J11);
    const float P00 = dot3c(T00, c[0], T01, c[1], T02, c[2]), P10 = dot3c(T00, c[1], T01, c[3], T02, c[4]),
                P20 = dot3c(T00, c[2], T01, c[4], T02, c[5]);
    const float P01 = dot3c(T10, c[0], T11, c[1], T12, c[2]), P11 = dot3c(T10, c[1], T11, c[3], T12, c[4]),
                P21 = dot3c(T10, c[2], T11, c[4], T12, c[5]);
    out3[0] = dot3c(T00, P00, T01, P10, T02, P20);
    out3[1] = dot3c(T00, P01, T01, P11, T02, P21);
    out3[2] = dot3c(T10, P01, T11, P11, T12, P21);
}
/* ndc2Pix in double, auxiliary.h:66-69 */
static float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

/* evaluate_opacity_factor, stopthepop_common.cuh:76-79 */
static float opacity_factor(float dx, float dy, float A, float B, float C) {
    const float q = fmaf(dx, A * dx, dy * (C * dy));
    return fmaf(dy, B * dx, q * 0.5f);
}
/* power of renderCUDA / HIER / all backward kernels, forward.cu:309 */
static float gaussian_power(float dx, float dy, float A, float B, float C) {
    const float q = fmaf(dx, dx * A, dy * (dy * C));
    return fmaf(q, -0.5f, -(dy * (dx * B)));
}
static float saturatef(float x) { return !(x > 0.0f) ? 0.0f : (x > 1.0f ? 1.0f : x); }
/* max_contrib_power_rect_gaussian_float<PW,PH>, stopthepop_common.cuh:130-174 */
static float max_contrib_power(int PW, float A, float B, float C, float mx_, float my_, float rminx, float rminy,
                               float rmaxx, float rmaxy, float* ox, float* oy) {
    const float xmd = rminx - mx_, ymd = rminy - my_;
    const float x_left = rminx > mx_ ? 1.0f : 0.0f, y_above = rminy > my_ ? 1.0f : 0.0f;
    const float nix = x_left + (mx_ > rmaxx ? 1.0f : 0.0f), niy = y_above + (my_ > rmaxy ? 1.0f : 0.0f);
    *ox = mx_; *oy = my_;
    if (!((niy + nix) > 0.0f)) return 0.0f;
    const float px = fmaf(rminx, x_left, rmaxx * (1.0f - x_left)), py = fmaf(rminy, y_above, rmaxy * (1.0f - y_above));
    const float dx = copysignf((float)PW, xmd), dy = copysignf((float)PW, ymd);
    const float diffx = mx_ - px, diffy = my_ - py;
    const float rcpx = 1.0f / (A * (float)(PW * PW)), rcpy = 1.0f / (C * (float)(PW * PW));
    const float txn = fmaf(diffy, B * dx, diffx * (A * dx)), tyn = fmaf(diffy, C * dy, diffx * (B * dy));
    const float tx = niy * saturatef(txn * rcpx), ty = nix * saturatef(tyn * rcpy);
    *ox = fmaf(dx, tx, px); *oy = fmaf(dy, ty, py);
    return opacity_factor(mx_ - *ox, my_ - *oy, A, B, C);
}

/* pix2world + computeViewRay, auxiliary.h:71-81, stopthepop_common.cuh:68-74 */
static void view_ray(const OrcInputs* in, float pxl, float pyl, float* d) {
    const float* m = in->inv_viewproj;
    const float nx = fmaf(pxl, 2.0f / (float)in->W, -1.0f), ny = fmaf(pyl, 2.0f / (float)in->H, -1.0f);
    const float pw = m[15] + fmaf(m[3], nx, m[7] * ny), pz = m[14] + fmaf(m[2], nx, m[6] * ny);
    const float py = m[13] + fmaf(m[1], nx, m[5] * ny), px = m[12] + fmaf(m[0], nx, m[4] * ny);
    const float rw = 1.0f / pw;
    const float vx = fmaf(px, rw, -in->campos[0]), vy = fmaf(py, rw, -in->campos[1]), vz = fmaf(pz, rw, -in->campos[2]);
    const float inv = 1.0f / sqrtf(fmaf(vz, vz, fmaf(vx, vx, vy * vy)));
    d[0] = vx * inv; d[1] = vy * inv; d[2] = vz * inv;
}
/* the same ray as compiled inside renderSortedFullCUDA (resorted_render.cuh:521-533): the pixel is the pair of loop
 * counters x (outer) / y (inner), the x products leave the inner loop before FMA contraction, so the unprojection is
 * m3 + fma(m1, ny, m0 * nx) -- one rounding away from view_ray; pinned at 4K against the reference build
 * (tests/test_gpu_matrix.py::test_full_sort_at_benchmark_size_matches_reference_build). */
static void view_ray_xloop(const OrcInputs* in, float pxl, float pyl, float* d) {
    const float* m = in->inv_viewproj;
    const float nx = fmaf(pxl, 2.0f / (float)in->W, -1.0f), ny = fmaf(pyl, 2.0f / (float)in->H, -1.0f);
    const float pw = m[15] + fmaf(m[7], ny, m[3] * nx), pz = m[14] + fmaf(m[6], ny, m[2] * nx);
    const float py = m[13] + fmaf(m[5], ny, m[1] * nx), px = m[12] + fmaf(m[4], ny, m[0] * nx);
    const float rw = 1.0f / pw;
    const float vx = fmaf(px, rw, -in->campos[0]), vy = fmaf(py, rw, -in->campos[1]), vz = fmaf(pz, rw, -in->campos[2]);
    const float inv = 1.0f / sqrtf(fmaf(vz, vz, fmaf(vx, vx, vy * vy)));
    d[0] = vx * inv; d[1] = vy * inv; d[2] = vz * inv;
}
/* depthAlongRay, stopthepop_common.cuh:43-55; ic = cov3D_inv row of 12 floats */
static void depth_parts(const float* ic, const float* d, float* num, float* rcp) {
    const float vx = dot3c(ic[0], d[0], ic[1], d[1], ic[2], d[2]);
    const float vy = dot3c(ic[1], d[0], ic[4], d[1], ic[5], d[2]);
    const float vz = dot3c(ic[2], d[0], ic[5], d[1], ic[6], d[2]);
    *num = dot3c(ic[8], d[0], ic[9], d[1], ic[10], d[2]);
    *rcp = 1.0f / fmaxf(0.00001f, dot3c(d[0], vx, d[1], vy, d[2], vz));
}
static float depth_along_ray(const float* ic, const float* d) {
    float n, r;
    depth_parts(ic, d, &n, &r);
    return n * r;
}

static void tile_rect(const OrcInputs* in, const float* p, const float* e, int* r) { /* getRect, auxiliary.h:91-101 */
    const int gx = (in->W + 15) / 16, gy = (in->H + 15) / 16;
#define CL(v, g) ((v) < 0 ? 0 : ((v) > (g) ? (g) : (v)))
    const float a = floorf((p[0] - e[0]) * 0.0625f), b = floorf((p[1] - e[1]) * 0.0625f);
    const float c = ceilf((p[0] + e[0]) * 0.0625f), d = ceilf((p[1] + e[1]) * 0.0625f);
    const float big = 2.0e9f;
    r[0] = CL((int)fminf(fmaxf(a, -big), big), gx); r[1] = CL((int)fminf(fmaxf(b, -big), big), gy);
    r[2] = CL((int)fminf(fmaxf(c, -big), big), gx); r[3] = CL((int)fminf(fmaxf(d, -big), big), gy);
#undef CL
}

static const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                              -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
/* computeColorFromSH, forward_common.h:20-70 */
static void sh_to_rgb(int deg, const float* sh, const float* mean, const float* cam, float* rgb, uint8_t* clamped) {
    float d[3] = {mean[0] - cam[0], mean[1] - cam[1], mean[2] - cam[2]};
    const float len = sqrtf(fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1])));
    const float x = d[0] / len, y = d[1] / len, z = d[2] / len;
    for (int c = 0; c < 3; ++c) {
        float r = SH_C0 * sh[c];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] + SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] +
                    SH_C2[3] * xz * sh[21 + c] + SH_C2[4] * (xx - yy) * sh[24 + c];
                if (deg > 2)
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                        SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
            }
        }
        r += 0.5f;
        clamped[c] = r < 0.0f;
        rgb[c] = fmaxf(r, 0.0f);
    }
}

static int requires_inv(const OrcSettings* s) { return s->sort_mode != 0 || s->sort_order == 2 || s->sort_order == 3; }

/* preprocessCUDA, forward.cu:68-229 */
static void preprocess(const OrcInputs* in, const OrcSettings* s, OrcState* st) {
    const int P = in->P, gx = (in->W + 15) / 16;
    const float fy = in->H / (2.0f * in->tan_fovy), fx = in->W / (2.0f * in->tan_fovx); /* rasterizer_impl.cu:251-252 */
    const float* vm = in->viewmatrix;
    const float* pm = in->projmatrix;
    (void)gx;
    for (int i = 0; i < P; ++i) {
        st->radii[i] = 0;
        st->tiles_touched[i] = 0;
        const float x = in->means3D[3 * i], y = in->means3D[3 * i + 1], z = in->means3D[3 * i + 2];
        float pv[3];
        pv[0] = vm[12] + fmaf(z, vm[8], fmaf(x, vm[0], y * vm[4]));
        pv[1] = vm[13] + fmaf(z, vm[9], fmaf(x, vm[1], y * vm[5]));
        pv[2] = vm[14] + fmaf(z, vm[10], fmaf(x, vm[2], y * vm[6]));
        if (pv[2] <= 0.2f) continue; /* in_frustum, auxiliary.h:223 */
        float c6[6], R[3][3];
        if (in->cov3D_precomp) memcpy(c6, in->cov3D_precomp + 6 * i, sizeof(c6));
        else {
            quat_to_rot(in->rotations + 4 * i, R);
            gram(R, in->scale_modifier * in->scales[3 * i], in->scale_modifier * in->scales[3 * i + 1],
                 in->scale_modifier * in->scales[3 * i + 2], c6);
            memcpy(st->cov3D + 6 * i, c6, sizeof(c6));
        }
        float cv[3];
        cov2d(pv, fx, fy, in->tan_fovx, in->tan_fovy, c6, vm, cv);
        const float ca = cv[0] + 0.3f, cc = cv[2] + 0.3f, cb = cv[1], bb = cb * cb; /* dilateCov2D, forward_common.h:108-131 */
        const float det = fmaf(ca, cc, -bb);
        float scaling = 1.0f;
        if (s->proper_ewa_scaling) scaling = sqrtf(fmaxf(0.000025f, fmaf(cv[0], cv[2], -bb) / det));
        if (det == 0.0f) continue;
        const float di = 1.0f / det;
        const float co[4] = {cc * di, cb * -di, ca * di, in->opacities[i] * scaling}; /* computeConicOpacity :133-144 */
        if (co[3] < ALPHA_THRESHOLD) continue;
        const float thr = logf(co[3] / ALPHA_THRESHOLD);
        float extent = 3.33f;
        if (s->tight_opacity_bounding) extent = (float)fmin(3.33, (double)sqrtf(thr + thr));
        const float mid = 0.5f * (ca + cc);
        const float lambda = mid + sqrtf(fmaxf(0.01f, fmaf(mid, mid, -det)));
        const float radius = extent * sqrtf(lambda);
        if (radius <= 0.0f) continue;
        const float hx = fmaf(x, pm[0], y * pm[4]) + fmaf(z, pm[8], pm[12]);
        const float hy = fmaf(x, pm[1], y * pm[5]) + fmaf(z, pm[9], pm[13]);
        const float hw = fmaf(x, pm[3], y * pm[7]) + fmaf(z, pm[11], pm[15]);
        const float pw = 1.0f / (hw + 0.0000001f);
        const float m2[2] = {ndc2pix(hx * pw, in->W), ndc2pix(hy * pw, in->H)};
        const float ext[2] = {fminf(s->rect_bounding ? extent * sqrtf(ca) : radius, radius),
                              fminf(s->rect_bounding ? extent * sqrtf(cc) : radius, radius)};
        int rc[4];
        tile_rect(in, m2, ext, rc);
        int tiles = (rc[2] - rc[0]) * (rc[3] - rc[1]);
        if (tiles == 0) continue;
        if (s->tile_based_culling) { /* computeTilebasedCullingTileCount, stopthepop_common.cuh:176-262 */
            int cnt = 0;
            for (int ty = rc[1]; ty < rc[3]; ++ty)
                for (int tx = rc[0]; tx < rc[2]; ++tx) {
                    float ox, oy;
                    const float p = max_contrib_power(15, co[0], co[1], co[2], m2[0], m2[1], (float)(tx * 16), (float)(ty * 16),
                                                      (float)(tx * 16 + 15), (float)(ty * 16 + 15), &ox, &oy);
                    cnt += p <= thr;
                }
            tiles = cnt;
            if (tiles == 0) continue;
        }
        if (!in->colors_precomp) sh_to_rgb(in->D, in->shs + (size_t)i * in->M * 3, in->means3D + 3 * i, in->campos, st->rgb + 3 * i, st->clamped + 3 * i);
        else memcpy(st->rgb + 3 * i, in->colors_precomp + 3 * i, 12);
        const float vx = in->campos[0] - x, vy = in->campos[1] - y, vz = in->campos[2] - z;
        if (requires_inv(s)) { /* forward.cu:208-220 */
            float ic[6];
            quat_to_rot(in->rotations + 4 * i, R);
            gram(R, 1.0f / (in->scale_modifier * fmaxf(1e-3f, in->scales[3 * i])), 1.0f / (in->scale_modifier * fmaxf(1e-3f, in->scales[3 * i + 1])),
                 1.0f / (in->scale_modifier * fmaxf(1e-3f, in->scales[3 * i + 2])), ic);
            float* o = st->cov3D_inv + 12 * i;
            o[0] = ic[0]; o[1] = ic[1]; o[2] = ic[2]; o[3] = 0; o[4] = ic[3]; o[5] = ic[4]; o[6] = ic[5]; o[7] = 0;
            o[8] = fmaf(-ic[2], vz, fmaf(-ic[1], vy, -(ic[0] * vx)));
            o[9] = fmaf(-ic[4], vz, fmaf(-ic[3], vy, -(ic[1] * vx)));
            o[10] = fmaf(-ic[5], vz, fmaf(-ic[4], vy, -(ic[2] * vx)));
            o[11] = 0;
        }
        st->depths[i] = s->sort_order == 0 ? pv[2] : sqrtf(fmaf(vz, vz, fmaf(vx, vx, vy * vy)));
        st->radii[i] = (int)ceilf(radius);
        st->rects2D[2 * i] = ext[0]; st->rects2D[2 * i + 1] = ext[1];
        st->means2D[2 * i] = m2[0]; st->means2D[2 * i + 1] = m2[1];
        memcpy(st->conic_opacity + 4 * i, co, 16);
        st->tiles_touched[i] = (uint32_t)tiles;
    }
}

/* duplicateWithKeysCUDA (forward.cu:25-65) / duplicateWithKeys_extended (stopthepop_common.cuh:324-621),
 * stable sort on key bits [0,32+bit) (rasterizer_impl.cu:344-352), identifyTileRanges (:133-158) */
typedef struct { uint64_t k; uint32_t v; uint32_t seq; } KV;
static uint64_t g_mask;
static int kv_cmp(const void* a, const void* b) {
    const KV *x = a, *y = b;
    const uint64_t kx = x->k & g_mask, ky = y->k & g_mask;
    if (kx != ky) return kx < ky ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static uint32_t higher_msb(uint32_t n) { /* rasterizer_impl.cu:37-52 */
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
    if (n >> msb) msb++;
    return msb;
}
static void binning(const OrcInputs* in, const OrcSettings* s, OrcState* st) {
    const int P = in->P, gx = (in->W + 15) / 16, gy = (in->H + 15) / 16;
    uint32_t acc = 0;
    for (int i = 0; i < P; ++i) { acc += st->tiles_touched[i]; st->point_offsets[i] = acc; } /* InclusiveSum :313 */
    st->R = (int)acc;
    KV* kv = (KV*)malloc(sizeof(KV) * (acc ? acc : 1));
    const int ptd = s->sort_order == 2 || s->sort_order == 3;
    for (int i = 0; i < P; ++i) {
        if (st->radii[i] <= 0) continue;
        uint32_t off = i ? st->point_offsets[i - 1] : 0, end = st->point_offsets[i];
        int rc[4];
        tile_rect(in, st->means2D + 2 * i, st->rects2D + 2 * i, rc);
        const float* co = st->conic_opacity + 4 * i;
        const float thr = logf(co[3] / ALPHA_THRESHOLD);
        for (int ty = rc[1]; ty < rc[3]; ++ty)
            for (int tx = rc[0]; tx < rc[2]; ++tx) {
                float ox = 0, oy = 0, power = 0, depth = st->depths[i];
                if (s->tile_based_culling || s->sort_order == 3)
                    power = max_contrib_power(15, co[0], co[1], co[2], st->means2D[2 * i], st->means2D[2 * i + 1], (float)(tx * 16),
                                              (float)(ty * 16), (float)(tx * 16 + 15), (float)(ty * 16 + 15), &ox, &oy);
                if (ptd) { /* stopthepop_common.cuh:439-449 */
                    float tp[2] = {ox, oy}, d[3], num, rcp;
                    if (s->sort_order == 2) { tp[0] = ((float)(tx * 16) + (float)(tx * 16 + 15)) * 0.5f; tp[1] = ((float)(ty * 16) + (float)(ty * 16 + 15)) * 0.5f; }
                    view_ray(in, tp[0], tp[1], d);
                    depth_parts(st->cov3D_inv + 12 * i, d, &num, &rcp);
                    depth = fmaxf(0.0f, fmaf(num, rcp, 8.0f));
                }
                if (s->tile_based_culling && !(power <= thr)) continue;
                if (off < end) {
                    uint32_t bits; memcpy(&bits, &depth, 4);
                    kv[off].k = ((uint64_t)(uint32_t)(ty * gx + tx) << 32) | bits; kv[off].v = (uint32_t)i; kv[off].seq = off;
                }
                ++off;
            }
        for (; off < end; ++off) { /* shortfall padding :504-508 */
            float fm = FLT_MAX; uint32_t bits; memcpy(&bits, &fm, 4);
            kv[off].k = ((uint64_t)0xFFFFFFFFu << 32) | bits; kv[off].v = 0xFFFFFFFFu; kv[off].seq = off;
        }
    }
    g_mask = (1ull << (32 + higher_msb((uint32_t)(gx * gy)))) - 1ull;
    qsort(kv, acc, sizeof(KV), kv_cmp);
    st->keys = (uint64_t*)malloc(8 * (acc ? acc : 1));
    st->point_list = (uint32_t*)malloc(4 * (acc ? acc : 1));
    for (uint32_t j = 0; j < acc; ++j) { st->keys[j] = kv[j].k; st->point_list[j] = kv[j].v; }
    free(kv);
    memset(st->ranges, 0, sizeof(uint32_t) * 2 * gx * gy);
    for (uint32_t j = 0; j < acc; ++j) {
        const uint32_t cur = (uint32_t)(st->keys[j] >> 32);
        const int valid = cur != 0xFFFFFFFFu;
        if (j == 0) { if (valid) st->ranges[2 * cur] = 0; }
        else {
            const uint32_t prev = (uint32_t)(st->keys[j - 1] >> 32);
            if (cur != prev) { if (prev != 0xFFFFFFFFu) st->ranges[2 * prev + 1] = j; if (valid) st->ranges[2 * cur] = j; }
        }
        if (j == acc - 1 && valid) st->ranges[2 * cur + 1] = acc;
    }
}

/* ---- GLOBAL render, forward.cu:234-366 ------------------------------------------------------------- */
/* ---- debug visualisation accumulators: accumSortingErrorDepth / outputDebugVis, stopthepop_common.cuh:264-307 ------
 * The reference compiles each render kernel a second time (ENABLE_DEBUG_VIZ) with these two calls inside the blend loop;
 * here the four render functions below carry them behind g_vis_type (0 = off, else the STP_DEBUG_* codes of
 * include/stp_rasterizer.h: 1 SortErrorOpacity, 2 SortErrorDistance, 3 GaussianCountPerTile, 4 Depth,
 * 5 GaussianCountPerPixel, 6 Transmittance).  orc_debug_visualisation() below returns the raw planes (value, T) that
 * render_debug_CUDA (forward.cu:674-714) then normalises and colour-maps (oracle/cpu_oracle.py:colormap).
 * GaussianCountPerPixel: the reference writes its loop counter `contributor` (list entries visited, including the ones
 * skipped), which depends on the kernel's loop structure; this restatement counts the Gaussians BLENDED, the definition
 * the CUDA path uses ("Deliberate differences"). */
static int g_vis_type = 0;
static float* g_vis_out = NULL; /* [2][H*W] */
typedef struct { float cur, acc; uint32_t blends; } Vis;
static void vis_init(Vis* v) { v->cur = -FLT_MAX; v->acc = 0.0f; v->blends = 0; } /* forward.cu:282, hierarchical_render.cuh:983 */
static void vis_accum(Vis* v, float depth, float alpha, float T) { /* T: transmittance BEFORE this blend */
    if ((g_vis_type == 1 || g_vis_type == 2) && depth <= v->cur) {
        if (g_vis_type == 1) v->acc += alpha;
        else v->acc += fabsf(v->cur - depth);
    } else if (g_vis_type == 4) {
        v->acc += depth * alpha * T;
    }
    v->cur = fmaxf(v->cur, depth);
    v->blends++;
}
static void vis_store(const OrcInputs* in, const Vis* v, int px, int py, float T, uint32_t range) {
    if (!g_vis_out || !(px < in->W && py < in->H)) return;
    const size_t N = (size_t)in->W * in->H, pid = (size_t)py * in->W + px;
    float out = 0.0f;
    switch (g_vis_type) {
        case 1: case 2: case 4: out = v->acc; break;
        case 3: out = (float)range; break;
        case 5: out = (float)v->blends; break;
        case 6: out = 1.0f - T; break;
    }
    g_vis_out[pid] = out;
    g_vis_out[N + pid] = T;
}

static void render_global(const OrcInputs* in, OrcState* st) {
    const int W = in->W, H = in->H, gx = (W + 15) / 16;
#pragma omp parallel for schedule(dynamic, 64)
    for (int pid = 0; pid < W * H; ++pid) {
        const int px = pid % W, py = pid / W;
        const uint32_t* rg = st->ranges + 2 * ((py / 16) * gx + px / 16);
        float T = 1.0f, C[3] = {0, 0, 0};
        uint32_t contributor = 0, last = 0;
        Vis vis; vis_init(&vis);
        for (uint32_t j = rg[0]; j < rg[1]; ++j) {
            ++contributor;
            const uint32_t id = st->point_list[j];
            const float* co = st->conic_opacity + 4 * id;
            const float dx = st->means2D[2 * id] - (float)px, dy = st->means2D[2 * id + 1] - (float)py;
            const float power = gaussian_power(dx, dy, co[0], co[1], co[2]);
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, co[3] * expf(power));
            if (alpha < ALPHA_THRESHOLD) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < T_THRESHOLD) break;
            for (int ch = 0; ch < 3; ++ch) C[ch] = fmaf(T, alpha * st->rgb[3 * id + ch], C[ch]);
            if (g_vis_type) { /* depth = distance of the centre from the camera, forward.cu:337-341 */
                const float ex = in->campos[0] - in->means3D[3 * id], ey = in->campos[1] - in->means3D[3 * id + 1],
                            ez = in->campos[2] - in->means3D[3 * id + 2];
                vis_accum(&vis, sqrtf(ex * ex + ey * ey + ez * ez), alpha, T);
            }
            T = test_T;
            last = contributor;
        }
        vis_store(in, &vis, px, py, T, rg[1] - rg[0]);
        st->final_T[pid] = T;
        st->n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ++ch) st->out_color[(size_t)ch * W * H + pid] = fmaf(T, in->bg[ch], C[ch]);
    }
}

/* ---- shared blend helpers for the re-sorting modes ----------------------------------------------------- */
typedef struct {
    float T, C[3];
    int active;
    /* backward */
    float T_final, g[3], final_color[3];
    Vis vis;
} Pix;

typedef struct {
    float *dmean2D, *dconic, *dopacity, *dcolor; /* accumulated in double precision-free float like the GPU */
    const float* dL_dpix;
    const float* pixel_colors;
} BwdCtx;

/* per-hit gradient terms, backward.cu:541-592 (shared by all render-bwd kernels, front-to-back form
 * of hierarchical_render.cuh:1098-1165 / resorted_render.cuh:316-384) */
static int blend_bwd_front_to_back(const OrcInputs* in, const OrcState* st, BwdCtx* b, Pix* p, int px, int py, int id, float G) {
    const float* co = st->conic_opacity + 4 * id;
    const float alpha = fminf(0.99f, co[3] * G);
    const float test_T = p->T * (1.0f - alpha);
    if (test_T < T_THRESHOLD) return 0;
    const float dx = st->means2D[2 * id] - (float)px, dy = st->means2D[2 * id + 1] - (float)py;
    const float dchannel = alpha * p->T;
    float dL_dalpha = 0.0f;
    for (int ch = 0; ch < 3; ++ch) {
        const float c = st->rgb[3 * id + ch];
        p->C[ch] += c * alpha * p->T;
        const float rest = (p->final_color[ch] - p->C[ch]) / test_T;
        dL_dalpha += (c - rest) * p->g[ch];
#pragma omp atomic
        b->dcolor[3 * id + ch] += dchannel * p->g[ch];
    }
    dL_dalpha *= p->T;
    const float bg_dot = in->bg[0] * p->g[0] + in->bg[1] * p->g[1] + in->bg[2] * p->g[2];
    dL_dalpha += (-p->T_final / (1.0f - alpha)) * bg_dot;
    const float dL_dG = co[3] * dL_dalpha, gdx = G * dx, gdy = G * dy;
    const float dGx = -gdx * co[0] - gdy * co[1], dGy = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
    b->dmean2D[3 * id] += dL_dG * dGx * (0.5f * in->W);
#pragma omp atomic
    b->dmean2D[3 * id + 1] += dL_dG * dGy * (0.5f * in->H);
#pragma omp atomic
    b->dconic[4 * id] += -0.5f * gdx * dx * dL_dG;
#pragma omp atomic
    b->dconic[4 * id + 1] += -0.5f * gdx * dy * dL_dG;
#pragma omp atomic
    b->dconic[4 * id + 3] += -0.5f * gdy * dy * dL_dG;
#pragma omp atomic
    b->dopacity[id] += G * dL_dalpha;
    p->T = test_T;
    return 1;
}
static int blend_fwd(const OrcState* st, Pix* p, int id, float alpha, float depth) { /* hierarchical_render.cuh:992-1013 */
    const float test_T = p->T * (1.0f - alpha);
    if (test_T < T_THRESHOLD) return 0;
    for (int ch = 0; ch < 3; ++ch) p->C[ch] += st->rgb[3 * id + ch] * alpha * p->T;
    if (g_vis_type) vis_accum(&p->vis, depth, alpha, p->T); /* :1005-1008, resorted_render.cuh:105-108 */
    p->T = test_T;
    return 1;
}
static void pix_init(const OrcInputs* in, const OrcState* st, const BwdCtx* b, Pix* p, int px, int py) {
    const int W = in->W, H = in->H, inside = px < W && py < H;
    memset(p, 0, sizeof(*p));
    p->T = 1.0f;
    p->active = inside;
    vis_init(&p->vis);
    if (b && inside) {
        const int pid = py * W + px;
        p->T_final = st->final_T[pid];
        for (int ch = 0; ch < 3; ++ch) {
            p->g[ch] = b->dL_dpix[(size_t)ch * W * H + pid];
            p->final_color[ch] = b->pixel_colors[(size_t)ch * W * H + pid] - p->T_final * in->bg[ch];
        }
    }
}
static void pix_store(const OrcInputs* in, OrcState* st, const Pix* p, int px, int py, int write_ncontrib, uint32_t nc, uint32_t range) {
    const int W = in->W, H = in->H;
    if (!(px < W && py < H)) return;
    const int pid = py * W + px;
    vis_store(in, &p->vis, px, py, p->T, range); /* outputDebugVis */
    st->final_T[pid] = p->T;
    if (write_ncontrib) st->n_contrib[pid] = nc;
    for (int ch = 0; ch < 3; ++ch) st->out_color[(size_t)ch * W * H + pid] = p->C[ch] + p->T * in->bg[ch];
}

/* ---- HIER: sortGaussiansRayHierarchicaEvaluation, hierarchical_render.cuh:207-935 ------------------------ */
#define HEAD_MAX 16
#define MID_MAX 20
typedef struct { float d; int id; float store; } HeadE;
typedef struct {
    HeadE h[HEAD_MAX];
    int count;
    Pix pix;
    float ray[3];
    int px, py;
} HeadQ;
typedef struct { float d[MID_MAX]; int id[MID_MAX]; int count; float ray[3]; } MidQ;

/* batcherSort<32> compare-exchange network, hierarchical_render.cuh:158-192 (16 "threads", 32 values) */
static void batcher32(float* k, int* v) {
    for (uint32_t size = 2; size <= 32; size *= 2) {
        uint32_t stride = size / 2;
        for (uint32_t t = 0; t < 16; ++t) {
            const uint32_t pos = 2 * t - (t & (stride - 1));
            if (k[pos] > k[pos + stride]) { float a = k[pos]; k[pos] = k[pos + stride]; k[pos + stride] = a; int b = v[pos]; v[pos] = v[pos + stride]; v[pos + stride] = b; }
        }
        const uint32_t first = stride;
        for (stride = first / 2; stride > 0; stride /= 2)
            for (uint32_t t = 0; t < 16; ++t) {
                const uint32_t offset = t & (first - 1);
                const uint32_t pos = 2 * t - (t & (stride - 1));
                if (offset >= stride && k[pos - stride] > k[pos]) {
                    float a = k[pos - stride]; k[pos - stride] = k[pos]; k[pos] = a; int b = v[pos - stride]; v[pos - stride] = v[pos]; v[pos] = b;
                }
            }
    }
}

typedef struct {
    const OrcInputs* in; const OrcSettings* s; OrcState* st; BwdCtx* bwd;
    int HEAD, MID;
} HierCtx;

/* debug tap: blend list (id, alpha*T) of one pixel, used by tests to compare traversal order */
static int g_dbg_px = -1, g_dbg_py = -1, g_dbg_n = 0, g_dbg_cap = 0;
static int* g_dbg_ids = NULL;
static float* g_dbg_w = NULL;
static void head_blend_one(HierCtx* c, HeadQ* q) { /* blend_one, :386-417 */
    q->count--;
    if (!q->pix.active) return;
    int ok;
    const float T_before = q->pix.T;
    if (c->bwd) ok = blend_bwd_front_to_back(c->in, c->st, c->bwd, &q->pix, q->px, q->py, q->h[0].id, q->h[0].store);
    else ok = blend_fwd(c->st, &q->pix, q->h[0].id, q->h[0].store, q->h[0].d);
    if (!ok) { q->pix.active = 0; return; }
    if (q->px == g_dbg_px && q->py == g_dbg_py && g_dbg_n < g_dbg_cap) {
        g_dbg_ids[g_dbg_n] = q->h[0].id;
        g_dbg_w[g_dbg_n++] = T_before - q->pix.T; /* = alpha*T */
    }
    for (int i = 1; i < c->HEAD; ++i) q->h[i - 1] = q->h[i];
    q->h[c->HEAD - 1].d = FLT_MAX;
}
/* front4OneFromMid, :421-536: the 4 smallest entries of one quad's mid queue go to its 4 pixels */
static void front4(HierCtx* c, MidQ* m, HeadQ* hq /*4 pixels*/, int checkvalid) {
    const OrcState* st = c->st;
    int any = 0;
    for (int p = 0; p < 4; ++p) any |= hq[p].pix.active;
    if (any)
        for (int inner = 0; inner < 4; ++inner) {
            const int id = m->id[inner];
            for (int p = 0; p < 4; ++p) {
                HeadQ* q = &hq[p];
                if (q->count >= c->HEAD) head_blend_one(c, q);
                if (checkvalid && id == -1) continue;
                if (id < 0) continue; /* reference would read out of bounds; cannot happen */
                const float depth = depth_along_ray(st->cov3D_inv + 12 * id, q->ray);
                if (!q->pix.active || depth < 0.0f) continue;
                const float* co = st->conic_opacity + 4 * id;
                const float dx = st->means2D[2 * id] - (float)q->px, dy = st->means2D[2 * id + 1] - (float)q->py;
                const float power = gaussian_power(dx, dy, co[0], co[1], co[2]);
                if (power > 0.0f) continue;
                const float G = expf(power), alpha = fminf(0.99f, co[3] * G);
                if (alpha < ALPHA_THRESHOLD) continue;
                HeadE e = {depth, id, c->bwd ? G : alpha};
                for (int k = 0; k < c->HEAD; ++k)
                    if (e.d < q->h[k].d) { HeadE t = q->h[k]; q->h[k] = e; e = t; }
                q->count++;
            }
        }
    /* pop the four from the queue */
    for (int k = 4; k < c->MID; ++k) { m->d[k - 4] = m->d[k]; m->id[k - 4] = m->id[k]; }
    for (int k = c->MID - 4; k < c->MID; ++k) { m->d[k] = FLT_MAX; m->id[k] = -1; }
    m->count -= 4;
}
/* one group of 4 tail entries enters the mid queue of every quad, :566-677 */
static void mid_push_group(HierCtx* c, const float* tail_d, const int* tail_id, MidQ* mq, HeadQ* hq, int checkvalid) {
    (void)tail_d;
    for (int qd = 0; qd < 4; ++qd) {
        MidQ* m = &mq[qd];
        float nd[4]; int nid[4];
        for (int l = 0; l < 4; ++l) {
            const int id = tail_id[l];
            float d;
            if (id == -1) d = FLT_MAX; /* only while draining (checkvalid) */
            else d = depth_along_ray(c->st->cov3D_inv + 12 * id, m->ray);
            (void)checkvalid;
            nd[l] = d; nid[l] = id;
        }
        /* shflRankingLocal<4>: rank by (depth, lane) */
        float sd[4]; int sid[4];
        for (int l = 0; l < 4; ++l) {
            int rank = 0;
            for (int o = 0; o < 4; ++o) if (o != l && (nd[o] < nd[l] || (nd[o] == nd[l] && o < l))) ++rank;
            sd[rank] = nd[l]; sid[rank] = nid[l];
        }
        /* stable merge, resident first on ties (mergeSortRegToSmem :24-70 / mergeSortInto :73-127) */
        const int r = m->count; /* resident entries, sorted in m->d[0..r) */
        float od[MID_MAX + 4]; int oid[MID_MAX + 4];
        int a = 0, b = 0, n = 0;
        while (a < r || b < 4) {
            if (b >= 4 || (a < r && m->d[a] <= sd[b])) { od[n] = m->d[a]; oid[n] = m->id[a]; ++a; }
            else { od[n] = sd[b]; oid[n] = sid[b]; ++b; }
            ++n;
        }
        for (int k = 0; k < n; ++k) { m->d[k] = od[k]; m->id[k] = oid[k]; }
        m->count = n;
        if (m->count > c->MID - 4) front4(c, m, hq + 4 * qd, 0);
    }
}

static void render_hier_block(HierCtx* c, int tile_x, int tile_y, int bx, int by, uint32_t r0, uint32_t r1) {
    const OrcInputs* in = c->in; OrcState* st = c->st;
    const int cx = tile_x * 16 + 4 * bx, cy = tile_y * 16 + 4 * by;
    float tail_d[64]; int tail_id[64]; int tail_count = 0;
    float tail_ray[3];
    MidQ mq[4]; HeadQ hq[16];
    view_ray(in, (float)cx + 1.5f, (float)cy + 1.5f, tail_ray);
    for (int qd = 0; qd < 4; ++qd) {
        view_ray(in, (float)cx + (0.5f + 2 * (qd % 2)), (float)cy + (0.5f + 2 * (qd / 2)), mq[qd].ray);
        mq[qd].count = 0;
        for (int k = 0; k < MID_MAX; ++k) { mq[qd].d[k] = FLT_MAX; mq[qd].id[k] = -1; }
        for (int p = 0; p < 4; ++p) {
            HeadQ* q = &hq[4 * qd + p];
            q->px = cx + (qd % 2) * 2 + (p % 2); q->py = cy + (qd / 2) * 2 + (p / 2);
            q->count = 0;
            for (int k = 0; k < HEAD_MAX; ++k) { q->h[k].d = FLT_MAX; q->h[k].id = -1; q->h[k].store = 0; }
            pix_init(in, st, c->bwd, &q->pix, q->px, q->py);
            view_ray(in, (float)q->px, (float)q->py, q->ray);
        }
    }
    for (int k = 0; k < 64; ++k) { tail_d[k] = FLT_MAX; tail_id[k] = -1; }
    for (uint32_t progress = r0; progress < r1; progress += 32) {
        int any = 0;
        for (int p = 0; p < 16; ++p) any |= hq[p].pix.active;
        if (!any) break; /* per-block version of the per-warp early exit :692 (no observable difference) */
        float nd[32]; int nid[32]; int valid = 0;
        for (int l = 0; l < 32; ++l) {
            int id = -1;
            if (progress + l < r1) id = (int)st->point_list[progress + l];
            float d = FLT_MAX;
            if (id != -1) {
                int culled = 0;
                if (c->s->hier_culling) { /* :723-743 */
                    const float* co = st->conic_opacity + 4 * id;
                    float ox, oy;
                    const float pw = max_contrib_power(3, co[0], co[1], co[2], st->means2D[2 * id], st->means2D[2 * id + 1], (float)cx, (float)cy,
                                                       (float)cx + 3.0f, (float)cy + 3.0f, &ox, &oy);
                    culled = fminf(0.99f, co[3] * expf(-pw)) < ALPHA_THRESHOLD;
                }
                if (!culled) d = depth_along_ray(st->cov3D_inv + 12 * id, tail_ray);
            }
            nd[l] = d; nid[l] = d == FLT_MAX ? -1 : id;
            valid += nid[l] != -1;
        }
        batcher32(nd, nid);
        if (tail_count != 0) { /* mergeSortRegToSmem<32>, resident first on ties */
            float od[64]; int oid[64]; int a = 0, b = 0, n = 0;
            while (a < 32 || b < 32) {
                if (b >= 32 || (a < 32 && tail_d[a] <= nd[b])) { od[n] = tail_d[a]; oid[n] = tail_id[a]; ++a; }
                else { od[n] = nd[b]; oid[n] = nid[b]; ++b; }
                ++n;
            }
            memcpy(tail_d, od, sizeof(od)); memcpy(tail_id, oid, sizeof(oid));
        } else {
            memcpy(tail_d, nd, sizeof(nd)); memcpy(tail_id, nid, sizeof(nid));
            for (int k = 32; k < 64; ++k) { tail_d[k] = FLT_MAX; tail_id[k] = -1; }
        }
        tail_count += valid;
        for (int half = 0; half < 2; ++half)
            if (tail_count > 32) { /* :827-846 */
                for (int g = 0; g < 4; ++g) mid_push_group(c, tail_d + 4 * g, tail_id + 4 * g, mq, hq, 0);
                memmove(tail_d, tail_d + 16, sizeof(float) * 48); memmove(tail_id, tail_id + 16, sizeof(int) * 48);
                for (int k = 48; k < 64; ++k) { tail_d[k] = FLT_MAX; tail_id[k] = -1; }
                tail_count -= 16;
            }
    }
    int any = 0;
    for (int p = 0; p < 16; ++p) any |= hq[p].pix.active;
    if (any) { /* drain tail -> mid -> head, :855-925 */
        while (tail_count > 0) {
            mid_push_group(c, tail_d, tail_id, mq, hq, 1);
            memmove(tail_d, tail_d + 4, sizeof(float) * 60); memmove(tail_id, tail_id + 4, sizeof(int) * 60);
            for (int k = 60; k < 64; ++k) { tail_d[k] = FLT_MAX; tail_id[k] = -1; }
            tail_count -= tail_count < 4 ? tail_count : 4;
        }
        for (int qd = 0; qd < 4; ++qd)
            while (mq[qd].count > 0) front4(c, &mq[qd], hq + 4 * qd, 1);
        for (int p = 0; p < 16; ++p)
            while (hq[p].pix.active && hq[p].count > 0) head_blend_one(c, &hq[p]);
    }
    if (!c->bwd)
        for (int p = 0; p < 16; ++p) pix_store(in, st, &hq[p].pix, hq[p].px, hq[p].py, 0, 0, r1 - r0);
}
static void render_hier(const OrcInputs* in, const OrcSettings* s, OrcState* st, BwdCtx* bwd) {
    const int gx = (in->W + 15) / 16, gy = (in->H + 15) / 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int blk = 0; blk < gx * gy * 16; ++blk) {
        const int tile = blk / 16, sub = blk % 16;
        HierCtx c = {in, s, st, bwd, s->q_head, s->q_mid};
        render_hier_block(&c, tile % gx, tile / gx, sub % 4, sub / 4, st->ranges[2 * tile], st->ranges[2 * tile + 1]);
    }
}

/* ---- k-buffer: renderkBufferCUDA / renderkBufferBackwardCUDA, resorted_render.cuh:17-471 -------------------- */
static void render_kbuffer(const OrcInputs* in, const OrcSettings* s, OrcState* st, BwdCtx* bwd) {
    const int W = in->W, H = in->H, gx = (W + 15) / 16;
    static const int sizes[] = {1, 2, 4, 8, 12, 16, 20, 24}; /* forward.cu:410-425 */
    int Wn = 24;
    for (int k = 0; k < 8; ++k) if (s->q_head <= sizes[k]) { Wn = sizes[k]; break; }
#pragma omp parallel for schedule(dynamic, 64)
    for (int pid = 0; pid < W * H; ++pid) {
        const int px = pid % W, py = pid / W;
        const uint32_t* rg = st->ranges + 2 * ((py / 16) * gx + px / 16);
        Pix p; pix_init(in, st, bwd, &p, px, py);
        float ray[3]; view_ray(in, (float)px, (float)py, ray);
        HeadE q[24]; int num = 0; int done = 0; uint32_t contributor = 0;
        for (int k = 0; k < 24; ++k) { q[k].d = FLT_MAX; q[k].id = -1; q[k].store = 0; }
        for (uint32_t j = rg[0]; j < rg[1] && !done; ++j) {
            if (num == Wn) { /* blend_one */
                --num;
                const int ok = bwd ? blend_bwd_front_to_back(in, st, bwd, &p, px, py, q[0].id, q[0].store) : blend_fwd(st, &p, q[0].id, q[0].store, q[0].d);
                if (!ok) { done = 1; break; }
                for (int k = 1; k < Wn; ++k) q[k - 1] = q[k];
                q[Wn - 1].d = FLT_MAX;
            }
            ++contributor;
            const int id = (int)st->point_list[j];
            if (id < 0) break;
            const float* co = st->conic_opacity + 4 * id;
            const float dx = st->means2D[2 * id] - (float)px, dy = st->means2D[2 * id + 1] - (float)py;
            float G, alpha;
            if (!bwd) { /* forward spelling: positive factor, exp(-power), resorted_render.cuh:165-173 */
                const float pw = opacity_factor(dx, dy, co[0], co[1], co[2]);
                if (pw < 0.0f) continue;
                G = expf(-pw);
            } else {
                const float pw = gaussian_power(dx, dy, co[0], co[1], co[2]);
                if (pw > 0.0f) continue;
                G = expf(pw);
            }
            alpha = fminf(0.99f, co[3] * G);
            if (alpha < ALPHA_THRESHOLD) continue;
            const float depth = depth_along_ray(st->cov3D_inv + 12 * id, ray);
            if (depth < 0.0f) continue;
            HeadE e = {depth, id, bwd ? G : alpha};
            for (int k = 0; k < Wn; ++k) if (e.d < q[k].d) { HeadE t = q[k]; q[k] = e; e = t; }
            ++num;
        }
        if (!done)
            while (num > 0) {
                --num;
                const int ok = bwd ? blend_bwd_front_to_back(in, st, bwd, &p, px, py, q[0].id, q[0].store) : blend_fwd(st, &p, q[0].id, q[0].store, q[0].d);
                if (!ok) break;
                for (int k = 1; k < Wn; ++k) q[k - 1] = q[k];
                q[Wn - 1].d = FLT_MAX;
            }
        if (!bwd) pix_store(in, st, &p, px, py, 1, contributor, rg[1] - rg[0]);
    }
}

/* ---- full per-pixel sort: renderSortedFullCUDA, resorted_render.cuh:474-675 --------------------------------- */
typedef struct { float k; int v; int ord; } FS;
static int fs_cmp(const void* a, const void* b) {
    const FS *x = a, *y = b;
    if (x->k < y->k) return -1;
    if (x->k > y->k) return 1;
    return x->ord - y->ord;
}
/* bwd != NULL: NOT in the reference (it has no PPX_FULL backward, backward.cu:733-736).  Derived extension used to pin
 * the CUDA path's replay backward: the same per-pixel order as the forward pass, with the front-to-back gradient terms
 * the reference uses for its k-buffer / hierarchical backward (blend_bwd_front_to_back). */
static void render_full(const OrcInputs* in, OrcState* st, BwdCtx* bwd) {
    const int W = in->W, H = in->H, gx = (W + 15) / 16;
#pragma omp parallel for schedule(dynamic, 16)
    for (int pid = 0; pid < W * H; ++pid) {
        const int px = pid % W, py = pid / W;
        const uint32_t* rg = st->ranges + 2 * ((py / 16) * gx + px / 16);
        const int n = (int)(rg[1] - rg[0]), rounds = (n + 255) / 256;
        float ray[3]; view_ray_xloop(in, (float)px, (float)py, ray);
        FS win[1024]; /* blocked arrangement: thread t, item i -> win[4t+i] */
        for (int t = 0; t < 256; ++t)
            for (int i = 0; i < 3; ++i) {
                const int idx = i * 256 + t;
                FS* e = &win[4 * t + i + 1];
                if (idx < n) { e->v = (int)st->point_list[rg[0] + idx]; e->k = depth_along_ray(st->cov3D_inv + 12 * e->v, ray); }
                else { e->k = FLT_MAX; e->v = -1; }
            }
        float T = 1.0f, C[3] = {0, 0, 0}; uint32_t contributor = 0, last = 0; int done = 0, todo = n;
        Pix bp;
        Vis vis; vis_init(&vis);
        if (bwd) pix_init(in, st, bwd, &bp, px, py);
        for (int r = 0; r < rounds; ++r, todo -= 256) {
            for (int t = 0; t < 256; ++t) {
                const int idx = (r + 3) * 256 + t;
                FS* e = &win[4 * t];
                if (idx < n) { e->v = (int)st->point_list[rg[0] + idx]; e->k = depth_along_ray(st->cov3D_inv + 12 * e->v, ray); }
                else { e->k = FLT_MAX; e->v = -1; }
            }
            for (int k = 0; k < 1024; ++k) win[k].ord = k;
            qsort(win, 1024, sizeof(FS), fs_cmp);
            /* striped: rank r -> thread r%256 item r/256 ; emit ranks 0..255, keep the rest */
            FS keep[1024];
            for (int rk = 0; rk < 1024; ++rk) keep[4 * (rk % 256) + rk / 256] = win[rk];
            const int lim = todo < 256 ? todo : 256;
            for (int i = 0; i < lim && !done; ++i) {
                const FS* e = &win[i];
                if (e->v == -1) break;
                ++contributor;
                const int id = e->v;
                const float* co = st->conic_opacity + 4 * id;
                const float dx = st->means2D[2 * id] - (float)px, dy = st->means2D[2 * id + 1] - (float)py;
                const float pw = opacity_factor(dx, dy, co[0], co[1], co[2]);
                if (pw < 0.0f) continue;
                const float G = expf(-pw);
                const float alpha = fminf(0.99f, co[3] * G);
                if (alpha < ALPHA_THRESHOLD) continue;
                if (bwd) {
                    if (!blend_bwd_front_to_back(in, st, bwd, &bp, px, py, id, G)) done = 1;
                    continue;
                }
                const float test_T = T * (1.0f - alpha);
                if (test_T < T_THRESHOLD) { done = 1; continue; }
                for (int ch = 0; ch < 3; ++ch) C[ch] += st->rgb[3 * id + ch] * alpha * T;
                if (g_vis_type) vis_accum(&vis, e->k, alpha, T); /* resorted_render.cuh:645-648 */
                T = test_T;
                last = contributor;
            }
            memcpy(win, keep, sizeof(keep));
        }
        if (bwd) continue; /* the forward outputs stay as they are */
        vis_store(in, &vis, px, py, T, (uint32_t)n);
        st->final_T[pid] = T;
        st->n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ++ch) st->out_color[(size_t)ch * W * H + pid] = C[ch] + T * in->bg[ch];
    }
}

/* ---- GLOBAL backward render, backward.cu:437-595 -------------------------------------------------------------- */
static void render_global_bwd(const OrcInputs* in, const OrcState* st, BwdCtx* b) {
    const int W = in->W, H = in->H, gx = (W + 15) / 16;
#pragma omp parallel for schedule(dynamic, 64)
    for (int pid = 0; pid < W * H; ++pid) {
        const int px = pid % W, py = pid / W;
        const uint32_t* rg = st->ranges + 2 * ((py / 16) * gx + px / 16);
        const float T_final = st->final_T[pid];
        float T = T_final, g[3], acc[3] = {0, 0, 0}, lc[3] = {0, 0, 0}, last_alpha = 0;
        for (int ch = 0; ch < 3; ++ch) g[ch] = b->dL_dpix[(size_t)ch * W * H + pid];
        const float bg_dot = in->bg[0] * g[0] + in->bg[1] * g[1] + in->bg[2] * g[2];
        const uint32_t last = st->n_contrib[pid];
        uint32_t contributor = rg[1] - rg[0];
        for (uint32_t j = rg[1]; j-- > rg[0];) {
            --contributor;
            if (contributor >= last) continue;
            const uint32_t id = st->point_list[j];
            const float* co = st->conic_opacity + 4 * id;
            const float dx = st->means2D[2 * id] - (float)px, dy = st->means2D[2 * id + 1] - (float)py;
            const float power = gaussian_power(dx, dy, co[0], co[1], co[2]);
            if (power > 0.0f) continue;
            const float G = expf(power), alpha = fminf(0.99f, co[3] * G);
            if (alpha < ALPHA_THRESHOLD) continue;
            T = T / (1.0f - alpha);
            const float dchannel = alpha * T;
            float dL_dalpha = 0;
            for (int ch = 0; ch < 3; ++ch) {
                const float c = st->rgb[3 * id + ch];
                acc[ch] = last_alpha * lc[ch] + (1.0f - last_alpha) * acc[ch];
                lc[ch] = c;
                dL_dalpha += (c - acc[ch]) * g[ch];
#pragma omp atomic
                b->dcolor[3 * id + ch] += dchannel * g[ch];
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
            const float dL_dG = co[3] * dL_dalpha, gdx = G * dx, gdy = G * dy;
            const float dGx = -gdx * co[0] - gdy * co[1], dGy = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
            b->dmean2D[3 * id] += dL_dG * dGx * (0.5f * W);
#pragma omp atomic
            b->dmean2D[3 * id + 1] += dL_dG * dGy * (0.5f * H);
#pragma omp atomic
            b->dconic[4 * id] += -0.5f * gdx * dx * dL_dG;
#pragma omp atomic
            b->dconic[4 * id + 1] += -0.5f * gdx * dy * dL_dG;
#pragma omp atomic
            b->dconic[4 * id + 3] += -0.5f * gdy * dy * dL_dG;
#pragma omp atomic
            b->dopacity[id] += G * dL_dalpha;
        }
    }
}

/* ---- per-Gaussian backward: computeCov2DCUDA (backward.cu:146-312), preprocessCUDA bwd (:384-434),
 * computeColorFromSH bwd (:22-141), computeCov3D bwd (:316-379) --------------------------------------------------- */
static void preprocess_bwd(const OrcInputs* in, const OrcSettings* s, const OrcState* st, const float* dmean2D, const float* dconic,
                           float* dopacity, float* dcolor, float* dmean3D, float* dcov3D, float* dsh, float* dscale, float* drot) {
    const float* vm = in->viewmatrix; const float* pj = in->projmatrix;
    const float h_y = in->H / (2.0f * in->tan_fovy), h_x = in->W / (2.0f * in->tan_fovx);
    for (int i = 0; i < in->P; ++i) {
        if (!(st->radii[i] > 0)) continue;
        const float* mean = in->means3D + 3 * i;
        const float* c3 = in->cov3D_precomp ? in->cov3D_precomp + 6 * i : st->cov3D + 6 * i;
        const float dcx = dconic[4 * i], dcy = dconic[4 * i + 1], dcz = dconic[4 * i + 3];
        float t[3] = {vm[0] * mean[0] + vm[4] * mean[1] + vm[8] * mean[2] + vm[12], vm[1] * mean[0] + vm[5] * mean[1] + vm[9] * mean[2] + vm[13],
                      vm[2] * mean[0] + vm[6] * mean[1] + vm[10] * mean[2] + vm[14]};
        const float limx = 1.3f * in->tan_fovx, limy = 1.3f * in->tan_fovy, txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2]; t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f, yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = h_x / t[2], J02 = -(h_x * t[0]) / (t[2] * t[2]), J11 = h_y / t[2], J12 = -(h_y * t[1]) / (t[2] * t[2]);
        float T0[3], T1[3], V0[3], V1[3];
        for (int r = 0; r < 3; ++r) { T0[r] = vm[4 * r] * J00 + vm[4 * r + 2] * J02; T1[r] = vm[4 * r + 1] * J11 + vm[4 * r + 2] * J12; }
        const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        for (int r = 0; r < 3; ++r) { V0[r] = V[r][0] * T0[0] + V[r][1] * T0[1] + V[r][2] * T0[2]; V1[r] = V[r][0] * T1[0] + V[r][1] * T1[1] + V[r][2] * T1[2]; }
        float cxx = T0[0] * V0[0] + T0[1] * V0[1] + T0[2] * V0[2], cxy = T0[0] * V1[0] + T0[1] * V1[1] + T0[2] * V1[2],
              cyy = T1[0] * V1[0] + T1[1] * V1[1] + T1[2] * V1[2];
        const float det_orig = cxx * cyy - cxy * cxy;
        cxx += 0.3f; cyy += 0.3f;
        float dxx = 0, dxy = 0, dyy = 0;
        if (s->proper_ewa_scaling) {
            const float dp = cxx * cyy - cxy * cxy, ratio = det_orig / dp, hs = sqrtf(fmaxf(0.000025f, ratio));
            const float dop = dopacity[i], dh = dop * in->opacities[i];
            dopacity[i] = dop * hs;
            const float dir = ratio <= 0.000025f ? 0.f : dh / (2.f * hs), w = 0.3f;
            const float q = w * w + w * (cxx + cyy) + cxx * cyy - cxy * cxy, df = dir / (q * q);
            dxx = w * (w * cyy + cyy * cyy + cxy * cxy) * df; dyy = w * (w * cxx + cxx * cxx + cxy * cxy) * df; dxy = -2.f * w * cxy * (w + cxx + cyy) * df;
        }
        const float denom = cxx * cyy - cxy * cxy, d2i = 1.0f / (denom * denom + 0.0000001f);
        float dc[6] = {0, 0, 0, 0, 0, 0};
        if (d2i != 0) {
            dxx += d2i * (-cyy * cyy * dcx + 2 * cxy * cyy * dcy + (denom - cxx * cyy) * dcz);
            dyy += d2i * (-cxx * cxx * dcz + 2 * cxx * cxy * dcy + (denom - cxx * cyy) * dcx);
            dxy += d2i * 2 * (cxy * cyy * dcx - (denom + 2 * cxy * cxy) * dcy + cxx * cxy * dcz);
            dc[0] = T0[0] * T0[0] * dxx + T0[0] * T1[0] * dxy + T1[0] * T1[0] * dyy;
            dc[3] = T0[1] * T0[1] * dxx + T0[1] * T1[1] * dxy + T1[1] * T1[1] * dyy;
            dc[5] = T0[2] * T0[2] * dxx + T0[2] * T1[2] * dxy + T1[2] * T1[2] * dyy;
            dc[1] = 2 * T0[0] * T0[1] * dxx + (T0[0] * T1[1] + T0[1] * T1[0]) * dxy + 2 * T1[0] * T1[1] * dyy;
            dc[2] = 2 * T0[0] * T0[2] * dxx + (T0[0] * T1[2] + T0[2] * T1[0]) * dxy + 2 * T1[0] * T1[2] * dyy;
            dc[4] = 2 * T0[2] * T0[1] * dxx + (T0[1] * T1[2] + T0[2] * T1[1]) * dxy + 2 * T1[1] * T1[2] * dyy;
        }
        memcpy(dcov3D + 6 * i, dc, sizeof(dc));
        const float dT00 = 2 * V0[0] * dxx + V1[0] * dxy, dT01 = 2 * V0[1] * dxx + V1[1] * dxy, dT02 = 2 * V0[2] * dxx + V1[2] * dxy;
        const float dT10 = 2 * V1[0] * dyy + V0[0] * dxy, dT11 = 2 * V1[1] * dyy + V0[1] * dxy, dT12 = 2 * V1[2] * dyy + V0[2] * dxy;
        const float dJ00 = vm[0] * dT00 + vm[4] * dT01 + vm[8] * dT02, dJ02 = vm[2] * dT00 + vm[6] * dT01 + vm[10] * dT02;
        const float dJ11 = vm[1] * dT10 + vm[5] * dT11 + vm[9] * dT12, dJ12 = vm[2] * dT10 + vm[6] * dT11 + vm[10] * dT12;
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xg * -h_x * tz2 * dJ02, dty = yg * -h_y * tz2 * dJ12;
        const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t[0]) * tz3 * dJ02 + (2 * h_y * t[1]) * tz3 * dJ12;
        float dm[3] = {vm[0] * dtx + vm[1] * dty + vm[2] * dtz, vm[4] * dtx + vm[5] * dty + vm[6] * dtz, vm[8] * dtx + vm[9] * dty + vm[10] * dtz};
        {
            const float mw = 1.0f / (pj[3] * mean[0] + pj[7] * mean[1] + pj[11] * mean[2] + pj[15] + 0.0000001f);
            const float mul1 = (pj[0] * mean[0] + pj[4] * mean[1] + pj[8] * mean[2] + pj[12]) * mw * mw;
            const float mul2 = (pj[1] * mean[0] + pj[5] * mean[1] + pj[9] * mean[2] + pj[13]) * mw * mw;
            const float gx_ = dmean2D[3 * i], gy_ = dmean2D[3 * i + 1];
            dm[0] += (pj[0] * mw - pj[3] * mul1) * gx_ + (pj[1] * mw - pj[3] * mul2) * gy_;
            dm[1] += (pj[4] * mw - pj[7] * mul1) * gx_ + (pj[5] * mw - pj[7] * mul2) * gy_;
            dm[2] += (pj[8] * mw - pj[11] * mul1) * gx_ + (pj[9] * mw - pj[11] * mul2) * gy_;
        }
        if (in->shs) {
            const float dv[3] = {mean[0] - in->campos[0], mean[1] - in->campos[1], mean[2] - in->campos[2]};
            const float len = sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
            const float x = dv[0] / len, y = dv[1] / len, z = dv[2] / len;
            const float* sh = in->shs + (size_t)i * in->M * 3;
            float* o = dsh + (size_t)i * in->M * 3;
            float dRGB[3];
            for (int c = 0; c < 3; ++c) dRGB[c] = dcolor[3 * i + c] * (st->clamped[3 * i + c] ? 0.f : 1.f);
            float w[16]; memset(w, 0, sizeof(w));
            float ddx[3] = {0, 0, 0}, ddy[3] = {0, 0, 0}, ddz[3] = {0, 0, 0};
            int ncoef = 1;
            w[0] = SH_C0;
            if (in->D > 0) {
                ncoef = 4; w[1] = -SH_C1 * y; w[2] = SH_C1 * z; w[3] = -SH_C1 * x;
                for (int c = 0; c < 3; ++c) { ddx[c] = -SH_C1 * sh[9 + c]; ddy[c] = -SH_C1 * sh[3 + c]; ddz[c] = SH_C1 * sh[6 + c]; }
                if (in->D > 1) {
                    ncoef = 9;
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    w[4] = SH_C2[0] * xy; w[5] = SH_C2[1] * yz; w[6] = SH_C2[2] * (2.f * zz - xx - yy); w[7] = SH_C2[3] * xz; w[8] = SH_C2[4] * (xx - yy);
                    for (int c = 0; c < 3; ++c) {
                        ddx[c] += SH_C2[0] * y * sh[12 + c] + SH_C2[2] * 2.f * -x * sh[18 + c] + SH_C2[3] * z * sh[21 + c] + SH_C2[4] * 2.f * x * sh[24 + c];
                        ddy[c] += SH_C2[0] * x * sh[12 + c] + SH_C2[1] * z * sh[15 + c] + SH_C2[2] * 2.f * -y * sh[18 + c] + SH_C2[4] * 2.f * -y * sh[24 + c];
                        ddz[c] += SH_C2[1] * y * sh[15 + c] + SH_C2[2] * 2.f * 2.f * z * sh[18 + c] + SH_C2[3] * x * sh[21 + c];
                    }
                    if (in->D > 2) {
                        ncoef = 16;
                        w[9] = SH_C3[0] * y * (3.f * xx - yy); w[10] = SH_C3[1] * xy * z; w[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
                        w[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); w[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
                        w[14] = SH_C3[5] * z * (xx - yy); w[15] = SH_C3[6] * x * (xx - 3.f * yy);
                        for (int c = 0; c < 3; ++c) {
                            ddx[c] += SH_C3[0] * sh[27 + c] * 3.f * 2.f * xy + SH_C3[1] * sh[30 + c] * yz + SH_C3[2] * sh[33 + c] * -2.f * xy + SH_C3[3] * sh[36 + c] * -3.f * 2.f * xz +
                                      SH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * sh[42 + c] * 2.f * xz + SH_C3[6] * sh[45 + c] * 3.f * (xx - yy);
                            ddy[c] += SH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + SH_C3[1] * sh[30 + c] * xz + SH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                                      SH_C3[3] * sh[36 + c] * -3.f * 2.f * yz + SH_C3[4] * sh[39 + c] * -2.f * xy + SH_C3[5] * sh[42 + c] * -2.f * yz + SH_C3[6] * sh[45 + c] * -3.f * 2.f * xy;
                            ddz[c] += SH_C3[1] * sh[30 + c] * xy + SH_C3[2] * sh[33 + c] * 4.f * 2.f * yz + SH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) +
                                      SH_C3[4] * sh[39 + c] * 4.f * 2.f * xz + SH_C3[5] * sh[42 + c] * (xx - yy);
                        }
                    }
                }
            }
            for (int k = 0; k < ncoef; ++k) for (int c = 0; c < 3; ++c) o[3 * k + c] = w[k] * dRGB[c];
            const float dd[3] = {ddx[0] * dRGB[0] + ddx[1] * dRGB[1] + ddx[2] * dRGB[2], ddy[0] * dRGB[0] + ddy[1] * dRGB[1] + ddy[2] * dRGB[2],
                                 ddz[0] * dRGB[0] + ddz[1] * dRGB[1] + ddz[2] * dRGB[2]};
            const float sum2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2], is32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dm[0] += ((sum2 - dv[0] * dv[0]) * dd[0] - dv[1] * dv[0] * dd[1] - dv[2] * dv[0] * dd[2]) * is32;
            dm[1] += (-dv[0] * dv[1] * dd[0] + (sum2 - dv[1] * dv[1]) * dd[1] - dv[2] * dv[1] * dd[2]) * is32;
            dm[2] += (-dv[0] * dv[2] * dd[0] - dv[1] * dv[2] * dd[1] + (sum2 - dv[2] * dv[2]) * dd[2]) * is32;
        }
        memcpy(dmean3D + 3 * i, dm, sizeof(dm));
        if (in->scales) {
            const float* q = in->rotations + 4 * i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                   {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                   {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float sc[3] = {in->scale_modifier * in->scales[3 * i], in->scale_modifier * in->scales[3 * i + 1], in->scale_modifier * in->scales[3 * i + 2]};
            float M[3][3], dM[3][3], dMt[3][3];
            for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) M[c][rr] = sc[rr] * R[c][rr];
            const float dS[3][3] = {{dc[0], 0.5f * dc[1], 0.5f * dc[2]}, {0.5f * dc[1], dc[3], 0.5f * dc[4]}, {0.5f * dc[2], 0.5f * dc[4], dc[5]}};
            for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) dM[c][rr] = 2.0f * M[0][rr] * dS[c][0] + 2.0f * M[1][rr] * dS[c][1] + 2.0f * M[2][rr] * dS[c][2];
            for (int k = 0; k < 3; ++k) dscale[3 * i + k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
            for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) dMt[k][j] = dM[j][k] * sc[k];
            drot[4 * i] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            drot[4 * i + 1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            drot[4 * i + 2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            drot[4 * i + 3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* ---- public entry points (called through ctypes by oracle/cpu_oracle.py) ------------------------------------------ */
OrcState* orc_forward(const OrcInputs* in, const OrcSettings* s) { /* Rasterizer::forward, rasterizer_impl.cu:221-413 */
    OrcState* st = (OrcState*)calloc(1, sizeof(OrcState));
    const int P = in->P, N = in->W * in->H, tiles = ((in->W + 15) / 16) * ((in->H + 15) / 16);
    st->P = P; st->W = in->W; st->H = in->H; st->tiles = tiles;
    st->radii = calloc(P ? P : 1, 4); st->depths = calloc(P ? P : 1, 4); st->means2D = calloc(P ? P : 1, 8); st->rects2D = calloc(P ? P : 1, 8);
    st->conic_opacity = calloc(P ? P : 1, 16); st->rgb = calloc(P ? P : 1, 12); st->cov3D = calloc(P ? P : 1, 24); st->cov3D_inv = calloc(P ? P : 1, 48);
    st->clamped = calloc(P ? P : 1, 3); st->tiles_touched = calloc(P ? P : 1, 4); st->point_offsets = calloc(P ? P : 1, 4);
    st->ranges = calloc(tiles, 8); st->out_color = calloc(N, 12); st->final_T = calloc(N, 4); st->n_contrib = calloc(N, 4);
    preprocess(in, s, st);
    binning(in, s, st);
    switch (s->sort_mode) {
        case 0: render_global(in, st); break;
        case 1: render_full(in, st, NULL); break;
        case 2: render_kbuffer(in, s, st, NULL); break;
        default: render_hier(in, s, st, NULL); break;
    }
    return st;
}
/* the ENABLE_DEBUG_VIZ forward pass: raw[0..N) = the visualised quantity, raw[N..2N) = final T (outputDebugVis); the state
 * of the run (ranges, final_T, ...) is returned like orc_forward's.  Not re-entrant (g_vis_* are process globals). */
OrcState* orc_debug_visualisation(const OrcInputs* in, const OrcSettings* s, int type, float* raw) {
    g_vis_type = type; g_vis_out = raw;
    OrcState* st = orc_forward(in, s);
    g_vis_type = 0; g_vis_out = NULL;
    return st;
}
/* off: mirror the reference (PPX_FULL backward is an error); on: the derived extension of render_full above */
static int g_allow_full_backward_ext = 0;
void orc_allow_full_backward_ext(int on) { g_allow_full_backward_ext = on; }

/* Rasterizer::backward, rasterizer_impl.cu:417-526; all outputs zero-filled by the caller */
int orc_backward(const OrcInputs* in, const OrcSettings* s, OrcState* st, const float* pixel_colors, const float* dL_dpix, float* dmean2D,
                 float* dconic, float* dopacity, float* dcolor, float* dmean3D, float* dcov3D, float* dsh, float* dscale, float* drot) {
    BwdCtx b = {dmean2D, dconic, dopacity, dcolor, dL_dpix, pixel_colors};
    switch (s->sort_mode) {
        case 0: render_global_bwd(in, st, &b); break;
        case 1:
            if (!g_allow_full_backward_ext) return -2; /* "Backward not supported for full per-pixel sort", backward.cu:735 */
            render_full(in, st, &b);
            break;
        case 2: render_kbuffer(in, s, st, &b); break;
        default: render_hier(in, s, st, &b); break;
    }
    preprocess_bwd(in, s, st, dmean2D, dconic, dopacity, dcolor, dmean3D, dcov3D, dsh, dscale, drot);
    return 0;
}
/* debug: forward blend list of pixel (px,py) in HIER mode; returns the number of blends */
int orc_debug_hier_pixel(const OrcInputs* in, const OrcSettings* s, OrcState* st, int px, int py, int* ids, float* w, int cap) {
    const int gx = (in->W + 15) / 16, tile = (py / 16) * gx + px / 16;
    HierCtx c = {in, s, st, NULL, s->q_head, s->q_mid};
    g_dbg_px = px; g_dbg_py = py; g_dbg_n = 0; g_dbg_cap = cap; g_dbg_ids = ids; g_dbg_w = w;
    render_hier_block(&c, px / 16, py / 16, (px % 16) / 4, (py % 16) / 4, st->ranges[2 * tile], st->ranges[2 * tile + 1]);
    g_dbg_px = g_dbg_py = -1;
    return g_dbg_n;
}
void orc_free(OrcState* st) {
    if (!st) return;
    free(st->radii); free(st->depths); free(st->means2D); free(st->rects2D); free(st->conic_opacity); free(st->rgb); free(st->cov3D);
    free(st->cov3D_inv); free(st->clamped); free(st->tiles_touched); free(st->point_offsets); free(st->keys); free(st->point_list);
    free(st->ranges); free(st->out_color); free(st->final_T); free(st->n_contrib); free(st);
}
