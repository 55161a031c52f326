// stp_math.cuh -- rounding-pinned device arithmetic shared by all kernels.
//
// Bit-exact index buffers (point_list / ranges) require that every float that feeds an integer
// decision -- the depth bits of the sort key, floor/ceil of the tile rectangle, the opacity and
// transmittance thresholds -- is produced by the same SEQUENCE OF ROUNDED OPERATIONS as in the
// reference build.  The reference is compiled with nvcc's default -fmad=true, so which products are
// fused into an FMA is decided by the compiler (twice: NVVM and ptxas) and depends on inlining
// context; the ground truth is therefore the reference's SASS, which tools/sass_expr.py turns into
// expression DAGs.  Everything below is written with explicit round-to-nearest intrinsics
// (__fmul_rn / __fadd_rn / __fmaf_rn never contract) following those DAGs; each function cites the
// reference source line whose compiled arithmetic it reproduces.  DESIGN.md section "Arithmetic
// contract" lists the DAGs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stp {

#define STP_DEV __device__ __forceinline__

STP_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
STP_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
STP_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
STP_DEV float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
STP_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
STP_DEV float fsqrt(float a) { return __fsqrt_rn(a); }
STP_DEV float frcp(float a) { return __frcp_rn(a); }

constexpr float kAlphaThreshold = 1.0f / 255.0f;  // auxiliary.h:21-22
constexpr float kTThreshold = 0.0001f;            // auxiliary.h:23
constexpr float kNearPlane = 0.2f;                // auxiliary.h:223

struct Vec3 {
    float x, y, z;
};

// (a0*b0 + a1*b1) + a2*b2 as nvcc contracts it everywhere in the reference:
//   t = a1*b1 ; t = fma(a0,b0,t) ; t = fma(a2,b2,t)
STP_DEV float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return ffma(a2, b2, ffma(a0, b0, fmul(a1, b1)));
}

// p_view = viewmatrix(4x3, column-major) * (x,y,z,1)     [auxiliary.h:220-221, glm type_mat4x3.inl:469-478]
STP_DEV Vec3 view_transform(const float* __restrict__ vm, float x, float y, float z) {
    Vec3 p;
    p.x = fadd(vm[12], ffma(z, vm[8], ffma(x, vm[0], fmul(y, vm[4]))));
    p.y = fadd(vm[13], ffma(z, vm[9], ffma(x, vm[1], fmul(y, vm[5]))));
    p.z = fadd(vm[14], ffma(z, vm[10], ffma(x, vm[2], fmul(y, vm[6]))));
    return p;
}

// Rotation matrix entries of an UN-normalised quaternion (r,x,y,z), shared by computeCov3D
// (forward_common.h:149-183) and computeInvCov3D (stopthepop_common.cuh:13-41).
// R[c][r]: column c, row r (glm column-major constructor order).
struct Rot3 {
    float c0[3], c1[3], c2[3];
};
STP_DEV Rot3 quat_to_rot(float r, float x, float y, float z) {
    const float xz = fmul(x, z), zz = fmul(z, z), yy = fmul(y, y);
    const float rx = fmul(r, x), rz = fmul(r, z);
    const float xz_p_ry = ffma(r, y, xz), xz_m_ry = ffma(-r, y, xz);
    const float yz_m_rx = ffma(y, z, -rx), yz_p_rx = ffma(y, z, rx);
    const float xy_m_rz = ffma(x, y, -rz), xy_p_rz = ffma(x, y, rz);
    const float yy_zz = fadd(yy, zz), xx_yy = ffma(x, x, yy), xx_zz = ffma(x, x, zz);
    Rot3 R;
    R.c0[0] = fadd(-fadd(yy_zz, yy_zz), 1.0f);
    R.c0[1] = fadd(xy_m_rz, xy_m_rz);
    R.c0[2] = fadd(xz_p_ry, xz_p_ry);
    R.c1[0] = fadd(xy_p_rz, xy_p_rz);
    R.c1[1] = fadd(-fadd(xx_zz, xx_zz), 1.0f);
    R.c1[2] = fadd(yz_m_rx, yz_m_rx);
    R.c2[0] = fadd(xz_m_ry, xz_m_ry);
    R.c2[1] = fadd(yz_p_rx, yz_p_rx);
    R.c2[2] = fadd(-fadd(xx_yy, xx_yy), 1.0f);
    return R;
}

// out = transpose(M) * M with M = diag(s) * R, upper triangle (00,01,02,11,12,22).
// The zero entries of diag(s) contribute exact +-0 addends in the reference (glm mat3*mat3 is not
// sparsity aware), so M[c][r] == round(s_r * R[c][r]) for finite inputs.
STP_DEV void gram_scaled_rot(const Rot3& R, float s0, float s1, float s2, float* __restrict__ out6) {
    const float m00 = fmul(s0, R.c0[0]), m01 = fmul(s1, R.c0[1]), m02 = fmul(s2, R.c0[2]);
    const float m10 = fmul(s0, R.c1[0]), m11 = fmul(s1, R.c1[1]), m12 = fmul(s2, R.c1[2]);
    const float m20 = fmul(s0, R.c2[0]), m21 = fmul(s1, R.c2[1]), m22 = fmul(s2, R.c2[2]);
    out6[0] = dot3c(m00, m00, m01, m01, m02, m02);
    out6[1] = dot3c(m00, m10, m01, m11, m02, m12);
    out6[2] = dot3c(m00, m20, m01, m21, m02, m22);
    out6[3] = dot3c(m10, m10, m11, m11, m12, m12);
    out6[4] = dot3c(m10, m20, m11, m21, m12, m22);
    out6[5] = dot3c(m20, m20, m21, m21, m22, m22);
}

// EWA projection of the 3D covariance: upper-left 2x2 of transpose(T) * transpose(Vrk) * T,
// T = W*J  (forward_common.h:73-106).  Returns (cov00, cov01, cov11).
STP_DEV Vec3 project_cov2d(const Vec3& pv, float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                           const float* __restrict__ c, const float* __restrict__ vm) {
    const float limx = fmul(1.3f, tan_fovx), limy = fmul(1.3f, tan_fovy);
    const float txtz = fdiv(pv.x, pv.z), tytz = fdiv(pv.y, pv.z);
    const float tx = fmul(fminf(limx, fmaxf(-limx, txtz)), pv.z);
    const float ty = fmul(fminf(limy, fmaxf(-limy, tytz)), pv.z);
    const float tz2 = fmul(pv.z, pv.z);
    const float J00 = fdiv(focal_x, pv.z), J11 = fdiv(focal_y, pv.z);
    const float J02 = fdiv(-fmul(focal_x, tx), tz2), J12 = fdiv(-fmul(focal_y, ty), tz2);
    // T[c][r], rows r of W are vm[4r+0..2]
    const float T00 = ffma(vm[2], J02, fmul(vm[0], J00));
    const float T01 = ffma(vm[6], J02, fmul(vm[4], J00));
    const float T02 = ffma(vm[10], J02, fmul(vm[8], J00));
    const float T10 = ffma(vm[2], J12, fmul(vm[1], J11));
    const float T11 = ffma(vm[6], J12, fmul(vm[5], J11));
    const float T12 = ffma(vm[10], J12, fmul(vm[9], J11));
    // P = transpose(T) * Vrk
    const float P00 = dot3c(T00, c[0], T01, c[1], T02, c[2]);
    const float P10 = dot3c(T00, c[1], T01, c[3], T02, c[4]);
    const float P20 = dot3c(T00, c[2], T01, c[4], T02, c[5]);
    const float P01 = dot3c(T10, c[0], T11, c[1], T12, c[2]);
    const float P11 = dot3c(T10, c[1], T11, c[3], T12, c[4]);
    const float P21 = dot3c(T10, c[2], T11, c[4], T12, c[5]);
    Vec3 cov;
    cov.x = dot3c(T00, P00, T01, P10, T02, P20);
    cov.y = dot3c(T00, P01, T01, P11, T02, P21);
    cov.z = dot3c(T10, P01, T11, P11, T12, P21);
    return cov;
}

// mean2D = ndc2Pix(world2ndc(mean))   [auxiliary.h:66-69,83-90; glm type_mat4x4.inl:561-572]
// p_hom = (m0*x + m1*y) + (m2*z + m3);  ndc2Pix runs in double: fma(v + 1.0, S, -1.0) * 0.5
STP_DEV float ndc_to_pix(float v, int S) {
    const double t = __fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0);
    return __double2float_rn(__dmul_rn(t, 0.5));
}
STP_DEV float2 project_mean2d(const float* __restrict__ pm, float x, float y, float z, int W, int H) {
    const float hx = fadd(ffma(x, pm[0], fmul(y, pm[4])), ffma(z, pm[8], pm[12]));
    const float hy = fadd(ffma(x, pm[1], fmul(y, pm[5])), ffma(z, pm[9], pm[13]));
    const float hw = fadd(ffma(x, pm[3], fmul(y, pm[7])), ffma(z, pm[11], pm[15]));
    const float pw = fdiv(1.0f, fadd(hw, 0.0000001f));
    return make_float2(ndc_to_pix(fmul(hx, pw), W), ndc_to_pix(fmul(hy, pw), H));
}

// getRect (auxiliary.h:91-101) with an optional tile-row band [row0,row1) for multi-GPU sharding.
// Without a band row0=0,row1=grid_y and this is exactly the reference clamp.
struct TileRect {
    int x0, y0, x1, y1;
};
STP_DEV TileRect tile_rect(float2 p, float2 ext, int grid_x, int grid_y, int row0, int row1) {
    TileRect r;
    r.x0 = min(grid_x, max(0, (int)floorf(fmul(fsub(p.x, ext.x), 0.0625f))));
    r.y0 = min(grid_y, max(0, (int)floorf(fmul(fsub(p.y, ext.y), 0.0625f))));
    r.x1 = min(grid_x, max(0, (int)ceilf(fmul(fadd(p.x, ext.x), 0.0625f))));
    r.y1 = min(grid_y, max(0, (int)ceilf(fmul(fadd(p.y, ext.y), 0.0625f))));
    r.y0 = min(row1, max(row0, r.y0));
    r.y1 = min(row1, max(row0, r.y1));
    return r;
}

// 0.5*(A dx^2 + C dy^2) + B dx dy   (evaluate_opacity_factor, stopthepop_common.cuh:76-79)
STP_DEV float opacity_factor(float dx, float dy, float A, float B, float C) {
    const float q = ffma(dx, fmul(A, dx), fmul(dy, fmul(C, dy)));
    return ffma(dy, fmul(B, dx), fmul(q, 0.5f));
}

// -0.5*(A dx^2 + C dy^2) - B dx dy  as compiled in renderCUDA fwd/bwd and the HIER kernels
// (forward.cu:309, backward.cu:529, hierarchical_render.cuh:487)
STP_DEV float gaussian_power(float dx, float dy, float A, float B, float C) {
    const float q = ffma(dx, fmul(dx, A), fmul(dy, fmul(dy, C)));
    return ffma(q, -0.5f, -fmul(dy, fmul(dx, B)));
}

// Largest contribution of a 2D Gaussian inside an axis-aligned pixel rectangle
// (max_contrib_power_rect_gaussian_float<PW,PH>, stopthepop_common.cuh:130-174).
// rect = [rmin, rmax], PW = rmax.x - rmin.x (15 for a tile, 3 for a 4x4 block).
// Returns the power at the maximising position and that position in (mx,my).
template <int PW, int PH>
STP_DEV float max_contrib_power(float A, float B, float C, float mean_x, float mean_y, float rmin_x, float rmin_y,
                                float rmax_x, float rmax_y, float& mx, float& my) {
    const float x_min_diff = fsub(rmin_x, mean_x);
    const float x_left = (rmin_x > mean_x) ? 1.0f : 0.0f;
    const float not_in_x = fadd(x_left, (mean_x > rmax_x) ? 1.0f : 0.0f);
    const float y_min_diff = fsub(rmin_y, mean_y);
    const float y_above = (rmin_y > mean_y) ? 1.0f : 0.0f;
    const float not_in_y = fadd(y_above, (mean_y > rmax_y) ? 1.0f : 0.0f);
    mx = mean_x;
    my = mean_y;
    float power = 0.0f;
    if (fadd(not_in_y, not_in_x) > 0.0f) {
        const float px = ffma(rmin_x, x_left, fmul(rmax_x, fsub(1.0f, x_left)));
        const float py = ffma(rmin_y, y_above, fmul(rmax_y, fsub(1.0f, y_above)));
        const float dx = copysignf((float)PW, x_min_diff);
        const float dy = copysignf((float)PH, y_min_diff);
        const float diffx = fsub(mean_x, px), diffy = fsub(mean_y, py);
        const float rcp_x = frcp(fmul(A, (float)(PW * PW)));
        const float rcp_y = frcp(fmul(C, (float)(PH * PH)));
        const float tx_num = ffma(diffy, fmul(B, dx), fmul(diffx, fmul(A, dx)));
        const float ty_num = ffma(diffy, fmul(C, dy), fmul(diffx, fmul(B, dy)));
        const float tx = fmul(not_in_y, __saturatef(fmul(tx_num, rcp_x)));
        const float ty = fmul(not_in_x, __saturatef(fmul(ty_num, rcp_y)));
        mx = ffma(dx, tx, px);
        my = ffma(dy, ty, py);
        power = opacity_factor(fsub(mean_x, mx), fsub(mean_y, my), A, B, C);
    }
    return power;
}

// Per-frame camera constants for ray generation (pix2world + computeViewRay,
// auxiliary.h:71-81, stopthepop_common.cuh:68-74).
struct RayCam {
    float i0[4], i1[4], i3[4];  // columns 0,1,3 of the inverse view-projection (glm column-major)
    float cx, cy, cz;           // camera position
    float two_over_w, two_over_h;
};
STP_DEV RayCam make_raycam(const float* __restrict__ ivp, const float* __restrict__ cam, int W, int H) {
    RayCam rc;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        rc.i0[k] = ivp[k];
        rc.i1[k] = ivp[4 + k];
        rc.i3[k] = ivp[12 + k];
    }
    rc.cx = cam[0];
    rc.cy = cam[1];
    rc.cz = cam[2];
    rc.two_over_w = fdiv(2.0f, (float)W);
    rc.two_over_h = fdiv(2.0f, (float)H);
    return rc;
}
// normalize(unproject(pix) - campos)
STP_DEV Vec3 view_ray(const RayCam& rc, float pix_x, float pix_y) {
    const float nx = ffma(pix_x, rc.two_over_w, -1.0f);
    const float ny = ffma(pix_y, rc.two_over_h, -1.0f);
    const float pw = fadd(rc.i3[3], ffma(rc.i0[3], nx, fmul(rc.i1[3], ny)));
    const float pz = fadd(rc.i3[2], ffma(rc.i0[2], nx, fmul(rc.i1[2], ny)));
    const float py = fadd(rc.i3[1], ffma(rc.i0[1], nx, fmul(rc.i1[1], ny)));
    const float px = fadd(rc.i3[0], ffma(rc.i0[0], nx, fmul(rc.i1[0], ny)));
    const float rw = frcp(pw);
    const float vx = ffma(px, rw, -rc.cx), vy = ffma(py, rw, -rc.cy), vz = ffma(pz, rw, -rc.cz);
    const float len2 = ffma(vz, vz, ffma(vx, vx, fmul(vy, vy)));
    const float inv = fdiv(1.0f, fsqrt(len2));
    Vec3 d;
    d.x = fmul(vx, inv);
    d.y = fmul(vy, inv);
    d.z = fmul(vz, inv);
    return d;
}

// The same ray as compiled inside renderSortedFullCUDA (resorted_render.cuh:521-533), where the pixel is not the thread's
// own but the pair of loop counters x (outer) / y (inner): the x-dependent products are loop invariant and leave the
// inner loop BEFORE the multiply-adds are contracted, so the unprojection is  i3 + fma(i1, ny, i0 * nx)  instead of
// i3 + fma(i0, nx, i1 * ny).  One rounding apart -- enough to flip the order of two Gaussians whose depths along the
// ray agree to a few ulps (a handful of pixels at 4K).
STP_DEV Vec3 view_ray_xloop(const RayCam& rc, float pix_x, float pix_y) {
    const float nx = ffma(pix_x, rc.two_over_w, -1.0f);
    const float ny = ffma(pix_y, rc.two_over_h, -1.0f);
    const float pw = fadd(rc.i3[3], ffma(rc.i1[3], ny, fmul(rc.i0[3], nx)));
    const float pz = fadd(rc.i3[2], ffma(rc.i1[2], ny, fmul(rc.i0[2], nx)));
    const float py = fadd(rc.i3[1], ffma(rc.i1[1], ny, fmul(rc.i0[1], nx)));
    const float px = fadd(rc.i3[0], ffma(rc.i1[0], ny, fmul(rc.i0[0], nx)));
    const float rw = frcp(pw);
    const float vx = ffma(px, rw, -rc.cx), vy = ffma(py, rw, -rc.cy), vz = ffma(pz, rw, -rc.cz);
    const float len2 = ffma(vz, vz, ffma(vx, vx, fmul(vy, vy)));
    const float inv = fdiv(1.0f, fsqrt(len2));
    Vec3 d;
    d.x = fmul(vx, inv);
    d.y = fmul(vy, inv);
    d.z = fmul(vz, inv);
    return d;
}

// depthAlongRay (stopthepop_common.cuh:43-55): numerator and reciprocal denominator kept separate
// because one caller fuses the final product with "+ 8" (per-tile depth key).
//   ic = {i00,i01,i02, i11,i12,i22},  u = Sigma^-1 (mu - o)
STP_DEV void depth_along_ray_parts(const float* __restrict__ ic, float ux, float uy, float uz, const Vec3& d,
                                   float& num, float& rcp_den) {
    const float vx = dot3c(ic[0], d.x, ic[1], d.y, ic[2], d.z);
    const float vy = dot3c(ic[1], d.x, ic[3], d.y, ic[4], d.z);
    const float vz = dot3c(ic[2], d.x, ic[4], d.y, ic[5], d.z);
    num = dot3c(ux, d.x, uy, d.y, uz, d.z);
    const float den = dot3c(d.x, vx, d.y, vy, d.z, vz);
    rcp_den = frcp(fmaxf(0.00001f, den));
}
STP_DEV float depth_along_ray(const float* __restrict__ ic, float ux, float uy, float uz, const Vec3& d) {
    float num, rcp;
    depth_along_ray_parts(ic, ux, uy, uz, d, num, rcp);
    return fmul(num, rcp);
}
// depth key of PER_TILE_DEPTH_* sort orders: max(0, depthAlongRay + 8) with the +8 fused
// (stopthepop_common.cuh:448, compiled as FFMA(num, rcp_den, 8))
STP_DEV float per_tile_depth_key(const float* __restrict__ ic, float ux, float uy, float uz, const Vec3& d) {
    float num, rcp;
    depth_along_ray_parts(ic, ux, uy, uz, d, num, rcp);
    return fmaxf(0.0f, ffma(num, rcp, 8.0f));
}


}  // namespace stp
