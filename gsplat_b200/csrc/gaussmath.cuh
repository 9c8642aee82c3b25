// gaussmath.cuh -- per-gaussian device math (row-major 3x3, quaternions wxyz).
//
// Compiled with -fmad=false: every operation is an individually rounded IEEE op, evaluated in a fixed
// order, so results are reproducible against the CPU oracle bit for bit.  These kernels are
// HBM-bound; the lost FMA throughput is irrelevant.
//
// Reference semantics: include/Utils.cuh:81-123 (posW2C, covarW2C), :228-347 (quaternion, covariance
// and their VJPs), :448-492 (inverse / blur VJPs), :567-690 (persp_proj and VJP),
// csrc/SphericalHarmonicsCUDA.cu:48-439 (SH basis and VJP), csrc/SphericalHarmonics.cuh:40-78 (view dir).
#pragma once
#include "common.cuh"

namespace gsb
{
struct M3
{
    float m[9];
};

__device__ __forceinline__ M3 mul(const M3 &A, const M3 &B)
{
    M3 C;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            C.m[i * 3 + j] = A.m[i * 3 + 0] * B.m[0 * 3 + j] + A.m[i * 3 + 1] * B.m[1 * 3 + j] + A.m[i * 3 + 2] * B.m[2 * 3 + j];
    return C;
}
__device__ __forceinline__ M3 mul_bt(const M3 &A, const M3 &B)
{ // A * B^T
    M3 C;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            C.m[i * 3 + j] = A.m[i * 3 + 0] * B.m[j * 3 + 0] + A.m[i * 3 + 1] * B.m[j * 3 + 1] + A.m[i * 3 + 2] * B.m[j * 3 + 2];
    return C;
}
__device__ __forceinline__ M3 mul_at(const M3 &A, const M3 &B)
{ // A^T * B
    M3 C;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            C.m[i * 3 + j] = A.m[0 * 3 + i] * B.m[0 * 3 + j] + A.m[1 * 3 + i] * B.m[1 * 3 + j] + A.m[2 * 3 + i] * B.m[2 * 3 + j];
    return C;
}

// natural log from exactly rounded ops only (see oracle gs_log): frexp, atanh series.
__device__ __forceinline__ float exact_log(float x)
{
    int e;
    float m = frexpf(x, &e);
    if(m < 0.70710678118654752f)
    {
        m = m * 2.f;
        e -= 1;
    }
    const float f = m - 1.f;
    const float s = f / (2.f + f);
    const float z = s * s;
    const float p
        = z * (0.33333333333333333f + z * (0.2f + z * (0.14285714285714285f + z * (0.11111111111111111f + z * 0.09090909090909091f))));
    const float two_s = 2.f * s;
    return (float)e * 0.69314718055994531f + (two_s + two_s * p);
}

__device__ __forceinline__ M3 quat_to_rotmat(const float *q)
{
    float w = q[0], x = q[1], y = q[2], z = q[3];
    const float inv_norm = 1.f / sqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float xy = x * y, xz = x * z, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    M3 R;
    R.m[0] = 1.f - 2.f * (y2 + z2);
    R.m[1] = 2.f * (xy - wz);
    R.m[2] = 2.f * (xz + wy);
    R.m[3] = 2.f * (xy + wz);
    R.m[4] = 1.f - 2.f * (x2 + z2);
    R.m[5] = 2.f * (yz - wx);
    R.m[6] = 2.f * (xz - wy);
    R.m[7] = 2.f * (yz + wx);
    R.m[8] = 1.f - 2.f * (x2 + y2);
    return R;
}

// accumulate dL/dq given G = dL/dR
__device__ __forceinline__ void quat_to_rotmat_vjp(const float *q, const M3 &G, float *v_q)
{
    float w = q[0], x = q[1], y = q[2], z = q[3];
    const float inv_norm = 1.f / sqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm, y *= inv_norm, z *= inv_norm, w *= inv_norm;
    const float *g = G.m;
    const float vw = 2.f * (x * (g[7] - g[5]) + y * (g[2] - g[6]) + z * (g[3] - g[1]));
    const float vx = 2.f * (-2.f * x * (g[4] + g[8]) + y * (g[3] + g[1]) + z * (g[6] + g[2]) + w * (g[7] - g[5]));
    const float vy = 2.f * (x * (g[3] + g[1]) - 2.f * y * (g[0] + g[8]) + z * (g[7] + g[5]) + w * (g[2] - g[6]));
    const float vz = 2.f * (x * (g[6] + g[2]) + y * (g[7] + g[5]) - 2.f * z * (g[0] + g[4]) + w * (g[3] - g[1]));
    const float dot = vw * w + vx * x + vy * y + vz * z;
    v_q[0] += (vw - dot * w) * inv_norm;
    v_q[1] += (vx - dot * x) * inv_norm;
    v_q[2] += (vy - dot * y) * inv_norm;
    v_q[3] += (vz - dot * z) * inv_norm;
}

// Sigma = (R S)(R S)^T ; INV: Sigma^-1 = (R S^-1)(R S^-1)^T
template<bool INV>
__device__ __forceinline__ M3 quat_scale_to_sym(const float *q, const float *s)
{
    const M3 R = quat_to_rotmat(q);
    M3 M;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            M.m[i * 3 + j] = R.m[i * 3 + j] * (INV ? (1.f / s[j]) : s[j]);
    return mul_bt(M, M);
}

template<bool INV>
__device__ __forceinline__ void quat_scale_sym_vjp(const float *q, const float *s, const M3 &v_sym, float *v_q, float *v_s)
{
    const M3 R = quat_to_rotmat(q);
    float d[3];
#pragma unroll
    for(int j = 0; j < 3; ++j)
        d[j] = INV ? (1.f / s[j]) : s[j];
    M3 M, Gs;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
        {
            M.m[i * 3 + j]  = R.m[i * 3 + j] * d[j];
            Gs.m[i * 3 + j] = v_sym.m[i * 3 + j] + v_sym.m[j * 3 + i];
        }
    const M3 v_M = mul(Gs, M);
    M3 v_R;
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            v_R.m[i * 3 + j] = v_M.m[i * 3 + j] * d[j];
    quat_to_rotmat_vjp(q, v_R, v_q);
#pragma unroll
    for(int j = 0; j < 3; ++j)
    {
        const float col = R.m[0 * 3 + j] * v_M.m[0 * 3 + j] + R.m[1 * 3 + j] * v_M.m[1 * 3 + j] + R.m[2 * 3 + j] * v_M.m[2 * 3 + j];
        v_s[j] += INV ? (-d[j] * d[j] * col) : col;
    }
}

// ---------------------------------------------------------------------------------------------
struct Cam
{
    M3 R;
    float t[3];
    float fx, fy, cx, cy;
};

__device__ __forceinline__ Cam load_cam(const float *vm, const float *K)
{
    Cam c;
    c.R.m[0] = vm[0], c.R.m[1] = vm[1], c.R.m[2] = vm[2];
    c.R.m[3] = vm[4], c.R.m[4] = vm[5], c.R.m[5] = vm[6];
    c.R.m[6] = vm[8], c.R.m[7] = vm[9], c.R.m[8] = vm[10];
    c.t[0] = vm[3], c.t[1] = vm[7], c.t[2] = vm[11];
    c.fx = K[0], c.fy = K[4], c.cx = K[2], c.cy = K[5];
    return c;
}

struct PerspJ
{
    float J00, J11, J02, J12, tx, ty, rz, rz2;
    bool x_in, y_in;
};

__device__ __forceinline__ PerspJ persp_jacobian(const float *pc, const Cam &c, uint32_t W, uint32_t H)
{
    PerspJ o;
    const float x = pc[0], y = pc[1], z = pc[2];
    const float tan_fovx  = 0.5f * (float)W / c.fx;
    const float tan_fovy  = 0.5f * (float)H / c.fy;
    const float lim_x_pos = ((float)W - c.cx) / c.fx + 0.3f * tan_fovx;
    const float lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    const float lim_y_pos = ((float)H - c.cy) / c.fy + 0.3f * tan_fovy;
    const float lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    o.rz                  = 1.f / z;
    o.rz2                 = o.rz * o.rz;
    const float xz = x * o.rz, yz = y * o.rz;
    float cxz = xz > -lim_x_neg ? xz : -lim_x_neg;
    cxz       = cxz < lim_x_pos ? cxz : lim_x_pos;
    float cyz = yz > -lim_y_neg ? yz : -lim_y_neg;
    cyz       = cyz < lim_y_pos ? cyz : lim_y_pos;
    o.tx      = z * cxz;
    o.ty      = z * cyz;
    o.J00     = c.fx * o.rz;
    o.J11     = c.fy * o.rz;
    o.J02     = -c.fx * o.tx * o.rz2;
    o.J12     = -c.fy * o.ty * o.rz2;
    o.x_in    = (xz <= lim_x_pos && xz >= -lim_x_neg);
    o.y_in    = (yz <= lim_y_pos && yz >= -lim_y_neg);
    return o;
}

// ---- orthographic / fisheye camera models (reference CameraModelType ids: 0 pinhole, 1 ortho, 2 fisheye;
// include/Utils.cuh:498-526 ortho_proj, :692-731 fisheye_proj).  Dense 2x3 Jacobian, row-major.
constexpr int kCamPinhole = 0, kCamOrtho = 1, kCamFisheye = 2;
constexpr float kFisheyeEps = 0.0000001f;

template<int CAM>
__device__ __forceinline__ void dense_jacobian(const float *pc, const Cam &c, float *J, float &m2x, float &m2y)
{
    const float x = pc[0], y = pc[1], z = pc[2];
    if constexpr(CAM == kCamOrtho)
    {
        J[0] = c.fx, J[1] = 0.f, J[2] = 0.f;
        J[3] = 0.f, J[4] = c.fy, J[5] = 0.f;
        m2x = c.fx * x + c.cx;
        m2y = c.fy * y + c.cy;
    }
    else
    {
        const float xy_len = sqrtf(x * x + y * y) + kFisheyeEps;
        const float theta  = atan2f(xy_len, z + kFisheyeEps);
        m2x                = x * c.fx * theta / xy_len + c.cx;
        m2y                = y * c.fy * theta / xy_len + c.cy;
        const float x2 = x * x + kFisheyeEps, y2 = y * y, xy = x * y;
        const float x2y2 = x2 + y2;
        const float inv  = 1.f / (x2y2 + z * z);
        const float b    = atan2f(xy_len, z) / xy_len / x2y2;
        const float a    = z * inv / x2y2;
        J[0] = c.fx * (x2 * a + y2 * b);
        J[1] = c.fx * xy * (a - b);
        J[2] = -c.fx * x * inv;
        J[3] = c.fy * xy * (a - b);
        J[4] = c.fy * (y2 * a + x2 * b);
        J[5] = -c.fy * y * inv;
    }
}

// Forward-mode dual number over (x, y, z): the fisheye Jacobian's own derivative is obtained by pushing
// duals through the very expression dense_jacobian evaluates (the reference ships a hand-expanded closed
// form, Utils.cuh:733-846; the CPU oracle restates that one, so the two derivations check each other).
struct D3
{
    float v, dx, dy, dz;
};
__device__ __forceinline__ D3 d3(float v, float dx, float dy, float dz) { return D3{v, dx, dy, dz}; }
__device__ __forceinline__ D3 operator+(const D3 &a, const D3 &b) { return d3(a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dz + b.dz); }
__device__ __forceinline__ D3 operator-(const D3 &a, const D3 &b) { return d3(a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dz - b.dz); }
__device__ __forceinline__ D3 operator+(const D3 &a, float s) { return d3(a.v + s, a.dx, a.dy, a.dz); }
__device__ __forceinline__ D3 operator*(const D3 &a, const D3 &b)
{
    return d3(a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy, a.dz * b.v + a.v * b.dz);
}
__device__ __forceinline__ D3 operator*(float s, const D3 &a) { return d3(s * a.v, s * a.dx, s * a.dy, s * a.dz); }
__device__ __forceinline__ D3 operator/(const D3 &a, const D3 &b)
{
    const float ib = 1.f / b.v, q = a.v * ib;
    return d3(q, (a.dx - q * b.dx) * ib, (a.dy - q * b.dy) * ib, (a.dz - q * b.dz) * ib);
}
__device__ __forceinline__ D3 d3_sqrt(const D3 &a)
{
    const float r = sqrtf(a.v), h = r > 0.f ? 0.5f / r : 0.f; // on the optical axis the length has no gradient
    return d3(r, h * a.dx, h * a.dy, h * a.dz);
}
__device__ __forceinline__ D3 d3_atan2(const D3 &p, const D3 &q)
{
    const float w = 1.f / (p.v * p.v + q.v * q.v);
    return d3(atan2f(p.v, q.v), (q.v * p.dx - p.v * q.dx) * w, (q.v * p.dy - p.v * q.dy) * w, (q.v * p.dz - p.v * q.dz) * w);
}

// v_pc contribution of v_J through the fisheye Jacobian:  sum_k  dJ[k]/d{x,y,z} * v_J[k]
__device__ __forceinline__ void fisheye_jacobian_vjp(const float *pc, const Cam &c, const float *v_J, float *v_pc)
{
    const D3 x = d3(pc[0], 1.f, 0.f, 0.f), y = d3(pc[1], 0.f, 1.f, 0.f), z = d3(pc[2], 0.f, 0.f, 1.f);
    const D3 len  = d3_sqrt(x * x + y * y) + kFisheyeEps;
    const D3 x2   = x * x + kFisheyeEps, y2 = y * y, xy = x * y;
    const D3 x2y2 = x2 + y2;
    const D3 r2   = x2y2 + z * z;
    const D3 one  = d3(1.f, 0.f, 0.f, 0.f);
    const D3 inv  = one / r2;
    const D3 b    = d3_atan2(len, z) / len / x2y2;
    const D3 a    = z * inv / x2y2;
    const D3 amb  = a - b;
    const D3 Jd[6] = {c.fx * (x2 * a + y2 * b), c.fx * (xy * amb), (-c.fx) * (x * inv),
                      c.fy * (xy * amb), c.fy * (y2 * a + x2 * b), (-c.fy) * (y * inv)};
#pragma unroll
    for(int k = 0; k < 6; ++k)
    {
        v_pc[0] += Jd[k].dx * v_J[k];
        v_pc[1] += Jd[k].dy * v_J[k];
        v_pc[2] += Jd[k].dz * v_J[k];
    }
}

struct Proj
{
    int rx, ry; // 0,0 when culled
    float mx, my, depth, ca, cb, cc, comp;
};

// Forward projection of one (camera, gaussian).  cov = world covariance.
// Reference: csrc/ProjectionEWA3DGSFused.cu:38-219.
template<int CAM = kCamPinhole>
__device__ __forceinline__ Proj project_one(
    const float *mean, const M3 &cov, const float *opacity, const Cam &cam, uint32_t W, uint32_t H, float eps2d,
    float near_plane, float far_plane, float radius_clip, bool comp_scales_opacity
)
{
    Proj o;
    o.rx = o.ry = 0;
    o.mx = o.my = o.depth = o.ca = o.cb = o.cc = o.comp = 0.f;
    float pc[3];
#pragma unroll
    for(int i = 0; i < 3; ++i)
        pc[i] = cam.R.m[i * 3 + 0] * mean[0] + cam.R.m[i * 3 + 1] * mean[1] + cam.R.m[i * 3 + 2] * mean[2] + cam.t[i];
    if(pc[2] < near_plane || pc[2] > far_plane)
        return o;
    const M3 covc = mul_bt(mul(cam.R, cov), cam.R);
    float c00, c01, c10, c11, m2x, m2y;
    if constexpr(CAM == kCamPinhole)
    {
        const PerspJ pj = persp_jacobian(pc, cam, W, H);
        float T2[6];
#pragma unroll
        for(int j = 0; j < 3; ++j)
        {
            T2[j]     = pj.J00 * covc.m[0 * 3 + j] + pj.J02 * covc.m[2 * 3 + j];
            T2[3 + j] = pj.J11 * covc.m[1 * 3 + j] + pj.J12 * covc.m[2 * 3 + j];
        }
        c00 = T2[0] * pj.J00 + T2[2] * pj.J02;
        c01 = T2[1] * pj.J11 + T2[2] * pj.J12;
        c10 = T2[3] * pj.J00 + T2[5] * pj.J02;
        c11 = T2[4] * pj.J11 + T2[5] * pj.J12;
        m2x = cam.fx * pc[0] * pj.rz + cam.cx;
        m2y = cam.fy * pc[1] * pj.rz + cam.cy;
    }
    else
    {
        float J[6], T2[6];
        dense_jacobian<CAM>(pc, cam, J, m2x, m2y);
#pragma unroll
        for(int i = 0; i < 2; ++i)
#pragma unroll
            for(int j = 0; j < 3; ++j)
                T2[i * 3 + j] = J[i * 3 + 0] * covc.m[0 * 3 + j] + J[i * 3 + 1] * covc.m[1 * 3 + j] + J[i * 3 + 2] * covc.m[2 * 3 + j];
        c00 = T2[0] * J[0] + T2[1] * J[1] + T2[2] * J[2];
        c01 = T2[0] * J[3] + T2[1] * J[4] + T2[2] * J[5];
        c10 = T2[3] * J[0] + T2[4] * J[1] + T2[5] * J[2];
        c11 = T2[3] * J[3] + T2[4] * J[4] + T2[5] * J[5];
    }

    const float det_orig = c00 * c11 - c01 * c10;
    c00 += eps2d;
    c11 += eps2d;
    const float det_blur = c00 * c11 - c01 * c10;
    const float ratio    = det_orig / det_blur;
    const float floor_   = kMinCompensation * kMinCompensation;
    const float compensation = sqrtf(ratio > floor_ ? ratio : floor_);
    if(!(det_blur > 0.f))
        return o;
    const float ood = 1.f / det_blur;

    float extend = kGaussianExtend;
    if(opacity != nullptr)
    {
        float op = *opacity;
        if(comp_scales_opacity)
            op *= compensation;
        if(op < kAlphaThreshold)
            return o;
        const float e2 = sqrtf(2.f * exact_log(op / kAlphaThreshold));
        extend         = e2 < extend ? e2 : extend;
    }
    const float rx = ceilf(extend * sqrtf(c00));
    const float ry = ceilf(extend * sqrtf(c11));
    if(rx <= radius_clip && ry <= radius_clip)
        return o;
    if(m2x + rx <= 0.f || m2x - rx >= (float)W || m2y + ry <= 0.f || m2y - ry >= (float)H)
        return o;
    o.rx = (int)rx, o.ry = (int)ry;
    o.mx = m2x, o.my = m2y, o.depth = pc[2];
    o.ca = c11 * ood, o.cb = -c01 * ood, o.cc = c00 * ood;
    o.comp = compensation;
    return o;
}

struct ProjGrad
{
    float v_mean[3]; // world
    M3 v_cov;        // world covariance (full 3x3, not symmetrised)
    float v_pc[3];   // camera-space mean gradient (for v_viewmats)
    M3 v_covc;
};

// VJP of project_one for a visible gaussian.  Reference: csrc/ProjectionEWA3DGSFused.cu:376-638.
template<int CAM = kCamPinhole>
__device__ __forceinline__ ProjGrad project_one_vjp(
    const float *mean, const M3 &cov, const Cam &cam, uint32_t W, uint32_t H, float eps2d, const float *conic,
    float vm2x, float vm2y, float v_depth, const float *v_conic, bool has_comp, float comp, float v_comp
)
{
    ProjGrad g;
    const float a = conic[0], bb = conic[1], cc = conic[2];
    const float P[4]  = {a, bb, bb, cc};
    const float vP[4] = {v_conic[0], v_conic[1] * 0.5f, v_conic[1] * 0.5f, v_conic[2]};
    float tmp[4], vS[4];
    tmp[0] = P[0] * vP[0] + P[1] * vP[2];
    tmp[1] = P[0] * vP[1] + P[1] * vP[3];
    tmp[2] = P[2] * vP[0] + P[3] * vP[2];
    tmp[3] = P[2] * vP[1] + P[3] * vP[3];
    vS[0]  = -(tmp[0] * P[0] + tmp[1] * P[2]);
    vS[1]  = -(tmp[0] * P[1] + tmp[1] * P[3]);
    vS[2]  = -(tmp[2] * P[0] + tmp[3] * P[2]);
    vS[3]  = -(tmp[2] * P[1] + tmp[3] * P[3]);
    if(has_comp)
    {
        const float det_conic = P[0] * P[3] - P[1] * P[2];
        const float v_sqr     = v_comp * 0.5f / (comp + 1e-6f);
        const float om        = 1.f - comp * comp;
        vS[0] += v_sqr * (om * P[0] - eps2d * det_conic);
        vS[1] += v_sqr * (om * P[1]);
        vS[2] += v_sqr * (om * P[2]);
        vS[3] += v_sqr * (om * P[3] - eps2d * det_conic);
    }
    float pc[3];
#pragma unroll
    for(int i = 0; i < 3; ++i)
        pc[i] = cam.R.m[i * 3 + 0] * mean[0] + cam.R.m[i * 3 + 1] * mean[1] + cam.R.m[i * 3 + 2] * mean[2] + cam.t[i];
    const M3 covc   = mul_bt(mul(cam.R, cov), cam.R);
    PerspJ pj;
    float J[6];
    if constexpr(CAM == kCamPinhole)
    {
        pj   = persp_jacobian(pc, cam, W, H);
        J[0] = pj.J00, J[1] = 0.f, J[2] = pj.J02;
        J[3] = 0.f, J[4] = pj.J11, J[5] = pj.J12;
    }
    else
    {
        float m2x_unused, m2y_unused;
        dense_jacobian<CAM>(pc, cam, J, m2x_unused, m2y_unused);
    }
    float JtG[6];
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 2; ++j)
            JtG[i * 2 + j] = J[0 * 3 + i] * vS[0 * 2 + j] + J[1 * 3 + i] * vS[1 * 2 + j];
#pragma unroll
    for(int i = 0; i < 3; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            g.v_covc.m[i * 3 + j] = JtG[i * 2 + 0] * J[0 * 3 + j] + JtG[i * 2 + 1] * J[1 * 3 + j];
    float GJ[6], GtJ[6], v_J[6];
#pragma unroll
    for(int i = 0; i < 2; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
        {
            GJ[i * 3 + j]  = vS[i * 2 + 0] * J[0 * 3 + j] + vS[i * 2 + 1] * J[1 * 3 + j];
            GtJ[i * 3 + j] = vS[0 * 2 + i] * J[0 * 3 + j] + vS[1 * 2 + i] * J[1 * 3 + j];
        }
#pragma unroll
    for(int i = 0; i < 2; ++i)
#pragma unroll
        for(int j = 0; j < 3; ++j)
            v_J[i * 3 + j]
                = (GJ[i * 3 + 0] * covc.m[j * 3 + 0] + GJ[i * 3 + 1] * covc.m[j * 3 + 1] + GJ[i * 3 + 2] * covc.m[j * 3 + 2])
                + (GtJ[i * 3 + 0] * covc.m[0 * 3 + j] + GtJ[i * 3 + 1] * covc.m[1 * 3 + j] + GtJ[i * 3 + 2] * covc.m[2 * 3 + j]);
    if constexpr(CAM == kCamPinhole)
    {
        const float x = pc[0], y = pc[1];
        const float rz = pj.rz, rz2 = pj.rz2, rz3 = rz2 * rz;
        const float fx = cam.fx, fy = cam.fy;
        g.v_pc[0] = fx * rz * vm2x;
        g.v_pc[1] = fy * rz * vm2y;
        g.v_pc[2] = -(fx * x * vm2x + fy * y * vm2y) * rz2;
        if(pj.x_in)
            g.v_pc[0] += -fx * rz2 * v_J[2];
        else
            g.v_pc[2] += -fx * rz3 * v_J[2] * pj.tx;
        if(pj.y_in)
            g.v_pc[1] += -fy * rz2 * v_J[5];
        else
            g.v_pc[2] += -fy * rz3 * v_J[5] * pj.ty;
        g.v_pc[2] += -fx * rz2 * v_J[0] - fy * rz2 * v_J[4] + 2.f * fx * pj.tx * rz3 * v_J[2] + 2.f * fy * pj.ty * rz3 * v_J[5];
    }
    else
    {
#pragma unroll
        for(int j = 0; j < 3; ++j)
            g.v_pc[j] = J[0 * 3 + j] * vm2x + J[1 * 3 + j] * vm2y;
        if constexpr(CAM == kCamFisheye)
            fisheye_jacobian_vjp(pc, cam, v_J, g.v_pc);
    }
    g.v_pc[2] += v_depth;
#pragma unroll
    for(int j = 0; j < 3; ++j)
        g.v_mean[j] = cam.R.m[0 * 3 + j] * g.v_pc[0] + cam.R.m[1 * 3 + j] * g.v_pc[1] + cam.R.m[2 * 3 + j] * g.v_pc[2];
    g.v_cov = mul(mul_at(cam.R, g.v_covc), cam.R);
    return g;
}

// ---------------------------------------------------------------------------------------------
// Spherical harmonics basis with optional gradient (dual numbers over x,y,z).
template<bool GRAD>
struct Dual
{
    float v, x, y, z;
};
template<>
struct Dual<false>
{
    float v;
};

template<bool G>
__device__ __forceinline__ Dual<G> dmk(float v, float x, float y, float z)
{
    Dual<G> r;
    r.v = v;
    if constexpr(G)
        r.x = x, r.y = y, r.z = z;
    return r;
}
template<bool G>
__device__ __forceinline__ Dual<G> operator*(const Dual<G> &a, const Dual<G> &b)
{
    Dual<G> r;
    r.v = a.v * b.v;
    if constexpr(G)
    {
        r.x = a.v * b.x + a.x * b.v;
        r.y = a.v * b.y + a.y * b.v;
        r.z = a.v * b.z + a.z * b.v;
    }
    return r;
}
template<bool G>
__device__ __forceinline__ Dual<G> operator+(const Dual<G> &a, const Dual<G> &b)
{
    Dual<G> r;
    r.v = a.v + b.v;
    if constexpr(G)
        r.x = a.x + b.x, r.y = a.y + b.y, r.z = a.z + b.z;
    return r;
}
template<bool G>
__device__ __forceinline__ Dual<G> operator-(const Dual<G> &a, const Dual<G> &b)
{
    Dual<G> r;
    r.v = a.v - b.v;
    if constexpr(G)
        r.x = a.x - b.x, r.y = a.y - b.y, r.z = a.z - b.z;
    return r;
}
template<bool G>
__device__ __forceinline__ Dual<G> sc(float s, const Dual<G> &a)
{
    Dual<G> r;
    r.v = s * a.v;
    if constexpr(G)
        r.x = s * a.x, r.y = s * a.y, r.z = s * a.z;
    return r;
}
template<bool G>
__device__ __forceinline__ Dual<G> sadd(float s, const Dual<G> &a, float c)
{
    Dual<G> r = sc(s, a);
    r.v       = r.v + c;
    return r;
}

// Visits Y[k], k = 0 .. (DEG+1)^2 - 1 in increasing k, at unit direction (ux,uy,uz): f(k, Y_k).
// A visitor (instead of an array) keeps the 16-25 dual numbers out of the register file: each basis
// function is consumed as soon as it is formed.  Constants: Sloan, JCGT 2013.
template<int DEG, bool G, class F>
__device__ __forceinline__ void sh_visit(float ux, float uy, float uz, F &&f)
{
    const Dual<G> x = dmk<G>(ux, 1.f, 0.f, 0.f), y = dmk<G>(uy, 0.f, 1.f, 0.f), z = dmk<G>(uz, 0.f, 0.f, 1.f);
    f(0, dmk<G>(0.2820947917738781f, 0.f, 0.f, 0.f));
    if constexpr(DEG >= 1)
    {
        f(1, sc(-0.48860251190292f, y));
        f(2, sc(0.48860251190292f, z));
        f(3, sc(-0.48860251190292f, x));
    }
    if constexpr(DEG >= 2)
    {
        const Dual<G> z2 = z * z;
        const Dual<G> fTmp0B = sc(-1.092548430592079f, z);
        const Dual<G> fC1 = x * x - y * y;
        const Dual<G> fS1 = sc(2.f, x * y);
        const Dual<G> Y6 = sadd(0.9461746957575601f, z2, -0.3153915652525201f);
        f(4, sc(0.5462742152960395f, fS1));
        f(5, fTmp0B * y);
        f(6, Y6);
        f(7, fTmp0B * x);
        f(8, sc(0.5462742152960395f, fC1));
        if constexpr(DEG >= 3)
        {
            const Dual<G> fTmp0C = sadd(-2.285228997322329f, z2, 0.4570457994644658f);
            const Dual<G> fTmp1B = sc(1.445305721320277f, z);
            const Dual<G> fC2 = x * fC1 - y * fS1;
            const Dual<G> fS2 = x * fS1 + y * fC1;
            const Dual<G> Y12 = z * sadd(1.865881662950577f, z2, -1.119528997770346f);
            f(9, sc(-0.5900435899266435f, fS2));
            f(10, fTmp1B * fS1);
            f(11, fTmp0C * y);
            f(12, Y12);
            f(13, fTmp0C * x);
            f(14, fTmp1B * fC1);
            f(15, sc(-0.5900435899266435f, fC2));
            if constexpr(DEG >= 4)
            {
                const Dual<G> fTmp0D = z * sadd(-4.683325804901025f, z2, 2.007139630671868f);
                const Dual<G> fTmp1C = sadd(3.31161143515146f, z2, -0.47308734787878f);
                const Dual<G> fTmp2B = sc(-1.770130769779931f, z);
                const Dual<G> fC3 = x * fC2 - y * fS2;
                const Dual<G> fS3 = x * fS2 + y * fC2;
                f(16, sc(0.6258357354491763f, fS3));
                f(17, fTmp2B * fS2);
                f(18, fTmp1C * fS1);
                f(19, fTmp0D * y);
                f(20, sc(1.984313483298443f, z * Y12) - sc(1.006230589874905f, Y6));
                f(21, fTmp0D * x);
                f(22, fTmp1C * fC1);
                f(23, fTmp2B * fC2);
                f(24, sc(0.6258357354491763f, fC3));
            }
        }
    }
}

// array form (used where all values are needed at once)
template<int DEG, bool G>
__device__ __forceinline__ void sh_basis(float ux, float uy, float uz, Dual<G> *Y)
{
    sh_visit<DEG, G>(ux, uy, uz, [&](int k, const Dual<G> &v) { Y[k] = v; });
}

// dir = mean + R^T t  (= mean - camera position for a rigid viewmat)
__device__ __forceinline__ void sh_view_dir(const float *mean, const float *vm, float *dir)
{
    const float tx = vm[3], ty = vm[7], tz = vm[11];
    dir[0] = mean[0] + (vm[0] * tx + vm[4] * ty + vm[8] * tz);
    dir[1] = mean[1] + (vm[1] * tx + vm[5] * ty + vm[9] * tz);
    dir[2] = mean[2] + (vm[2] * tx + vm[6] * ty + vm[10] * tz);
}
} // namespace gsb
