// rowrec.cuh -- the per-gaussian part of a compositing record: conservative extents for the block-level culling and the
// pre-scaled conic.  Shared by the pack pass (raster.cu: one evaluation per intersection) and the fused projection
// (pergauss.cu: one evaluation per visible gaussian, written as a 64-byte row record that the pack pass then only gathers).
#pragma once

#include "common.cuh"

namespace gsb
{
// geom stream: (a, c) * -0.5 log2(e), b * -log2(e)  ->  -sigma * log2(e) = A dx^2 + C dy^2 + B dx dy
constexpr float kLog2e        = 1.4426950408889634f;
constexpr float kConicScaleAC = -0.5f * kLog2e, kConicScaleB = -kLog2e;
constexpr float kConicUnscaleAC = 1.0f / kConicScaleAC, kConicUnscaleB = 1.0f / kConicScaleB;

#ifdef __CUDACC__
// cull = {mx, my, ex, ey}, axis = {ux, uy, lu, lv}, geom = {a', b', c', opacity} of one gaussian (see RecordStreams).
__device__ __forceinline__ void row_record(
    const float mx, const float my, const float a, const float b, const float c, const float op, float4 &cull, float4 &axis,
    float4 &geom
)
{
    // conservative half extents of {alpha >= 1/255}: |dx| <= sqrt(t c / det), t = 2 ln(255 op)
    // and of the box oriented along the ellipse axes (unit u = eigenvector of the larger eigenvalue l1 of
    // [[a,b],[b,c]], v = (-uy, ux)): |d.u| <= sqrt(t / l1), |d.v| <= sqrt(t / l2).  All bounds inflated.
    float ex, ey, ux = 1.f, uy = 0.f, lu = 1e30f, lv = 1e30f;
    const float det = a * c - b * b;
    if(!(op >= kAlphaThreshold))
    {
        ex = ey = -1e30f; // can never reach the alpha threshold (NaN opacity lands here too)
    }
    else if(!(det > 0.f) || !(a > 0.f) || !(c > 0.f) || !isfinite(det))
    {
        ex = ey = 1e30f; // not a proper ellipse: never cull, let the exact test decide
    }
    else
    {
        const float t = 2.f * logf(op * 255.f) * 1.0001f + 1e-4f;
        ex            = sqrtf(t * c / det) * 1.0001f + 0.01f;
        ey            = sqrtf(t * a / det) * 1.0001f + 0.01f;
        if(!isfinite(ex) || !isfinite(ey))
            ex = ey = 1e30f;
        const float hd = 0.5f * (a - c);
        const float l1 = 0.5f * (a + c) + sqrtf(hd * hd + b * b);
        const float l2 = det / l1;
        float vx1 = b, vy1 = l1 - a, vx2 = l1 - c, vy2 = b;
        if(vx2 * vx2 + vy2 * vy2 > vx1 * vx1 + vy1 * vy1)
            vx1 = vx2, vy1 = vy2;
        const float nn = vx1 * vx1 + vy1 * vy1;
        if(nn > 1e-30f && l2 > 0.f && isfinite(l1))
        {
            const float inv = rsqrtf(nn);
            ux = vx1 * inv, uy = vy1 * inv;
            const float ru = sqrtf(t / l1), rv = sqrtf(t / l2);
            // an axis direction error delta (~1e-6 rad) shifts the far end of the other axis by delta * length
            lu = ru * 1.0001f + 0.01f + 4e-6f * rv;
            lv = rv * 1.0001f + 0.01f + 4e-6f * ru;
            if(!isfinite(lu) || !isfinite(lv))
                lu = lv = 1e30f;
        }
    }
    cull = make_float4(mx, my, ex, ey);
    axis = make_float4(ux, uy, lu, lv);
    // the conic is stored pre-multiplied so that the exponent of 2 falls out of three FMAs: vis = 2^(A dx^2 + C dy^2 + B dx dy)
    geom = make_float4(a * kConicScaleAC, b * kConicScaleB, c * kConicScaleAC, op);
}
#endif
} // namespace gsb
