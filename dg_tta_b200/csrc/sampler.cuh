// Coordinate arithmetic and trilinear corner set-up of the affine sampler, shared by csrc/affine_sample.cu and the fused
// inverse-warp + consistency-loss kernels (csrc/consistency_loss.cu).  See affine_sample.cu for the torch semantics followed.
#pragma once
#include "common.cuh"

namespace dgtta {

struct SampleParams {
    const float *in;
    const float *theta;
    float *out;
    const float *bias;      // NULL, or [B]: out = sample(in - bias[b]) + bias[b]  (get_batch's min shift, torch_utils.py:58-62)
    int B, C, Di, Hi, Wi, Do, Ho, Wo;
    // launch-constant pieces of the base-grid arithmetic, computed on the host with the same IEEE operations:
    float step_w, step_h, step_d;   // 2 / (n - 1)
    float rcp_w, rcp_h, rcp_d;      // RN(1 / n)
    int exact_div;                  // some n > DIVC_MAX_N: use the IEEE division instead of the corrected reciprocal
};

// x / n, correctly rounded, without the division subroutine: q = RN(x * r), r = RN(1/n); residual e = fma(-q, n, x) is
// exact; RN(q + e * r) is the correctly rounded quotient (Markstein).  Checked exhaustively against IEEE division for
// every base-grid value of every n <= 4096 (tools/check_divc.py); larger n take __fdiv_rn.
constexpr int DIVC_MAX_N = 4096;

__device__ __forceinline__ float base_coord(int i, int n, float step, float rcp, int exact_div)
{
    if (n <= 1) return 0.f;
    const float v = (i < n / 2) ? __fmaf_rn(step, (float)i, -1.f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.f);
    const float m = __fmul_rn(v, (float)(n - 1));
    if (exact_div) return __fdiv_rn(m, (float)n);
    const float q = __fmul_rn(m, rcp);
    return __fmaf_rn(__fmaf_rn(-q, (float)n, m), rcp, q);
}

__device__ __forceinline__ float unnormalize(float g, int size)
{
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
}

__device__ __forceinline__ float clip_coord(float v, int size) { return fminf(fmaxf(v, 0.f), (float)(size - 1)); }

struct Coords {
    float ix, iy, iz;
};

// Block = 32 (w) x 8 (h) output voxels of one d-plane.  Every thread derives its own three base coordinates (a dozen
// FP32 instructions, no division subroutine, no shared-memory table, no barrier).
constexpr int SBX = 32, SBY = 8;
constexpr int SAMPLE_THREADS = SBX * SBY;

template <int PAD>
__device__ __forceinline__ Coords source_coords(const SampleParams &P, int b, int w, int h, int d)
{
    const float *th = P.theta + b * 12;
    const float xn = base_coord(min(w, P.Wo - 1), P.Wo, P.step_w, P.rcp_w, P.exact_div);
    const float yn = base_coord(min(h, P.Ho - 1), P.Ho, P.step_h, P.rcp_h, P.exact_div);
    const float zn = base_coord(d, P.Do, P.step_d, P.rcp_d, P.exact_div);
    float t[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) t[k] = __ldg(th + k);
    // base_grid @ theta^T: the K = 4 products accumulated in order, first one rounded, the rest fused (see header)
    const float gx = __fadd_rn(__fmaf_rn(zn, t[2], __fmaf_rn(yn, t[1], __fmul_rn(xn, t[0]))), t[3]);
    const float gy = __fadd_rn(__fmaf_rn(zn, t[6], __fmaf_rn(yn, t[5], __fmul_rn(xn, t[4]))), t[7]);
    const float gz = __fadd_rn(__fmaf_rn(zn, t[10], __fmaf_rn(yn, t[9], __fmul_rn(xn, t[8]))), t[11]);
    Coords c;
    c.ix = unnormalize(gx, P.Wi); c.iy = unnormalize(gy, P.Hi); c.iz = unnormalize(gz, P.Di);
    if (PAD == DGTTA_PAD_BORDER) { c.ix = clip_coord(c.ix, P.Wi); c.iy = clip_coord(c.iy, P.Hi); c.iz = clip_coord(c.iz, P.Di); }
    return c;
}

struct Corners {
    float wgt[8];
    int off[8];   // element offset inside one channel volume, -1 = outside (zeros padding)
};

__device__ __forceinline__ Corners trilinear_corners(const SampleParams &P, const Coords &c)
{
    const float fx = floorf(c.ix), fy = floorf(c.iy), fz = floorf(c.iz);
    const float tx = c.ix - fx, ty = c.iy - fy, tz = c.iz - fz;
    // clamp before the int conversion so that wild coordinates cannot overflow
    const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)P.Wi), y0 = (int)fminf(fmaxf(fy, -2.f), (float)P.Hi),
              z0 = (int)fminf(fmaxf(fz, -2.f), (float)P.Di);
    Corners q;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        const bool ok = xx >= 0 && xx < P.Wi && yy >= 0 && yy < P.Hi && zz >= 0 && zz < P.Di;
        q.wgt[k] = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
        q.off[k] = ok ? (zz * P.Hi + yy) * P.Wi + xx : -1;
    }
    return q;
}

// host side: launch-constant pieces of the base-grid arithmetic, computed with the same IEEE operations
inline SampleParams make_sample_params(const float *in, const float *theta, float *out, const float *bias, int B, int C, int Di,
                                       int Hi, int Wi, int Do, int Ho, int Wo)
{
    SampleParams P;
    P.in = in; P.theta = theta; P.out = out; P.bias = bias;
    P.B = B; P.C = C; P.Di = Di; P.Hi = Hi; P.Wi = Wi; P.Do = Do; P.Ho = Ho; P.Wo = Wo;
    // volatile: keep the host compiler from folding these into anything but one IEEE division each
    volatile float two = 2.f, one = 1.f;
    P.step_w = Wo > 1 ? two / (float)(Wo - 1) : 0.f;
    P.step_h = Ho > 1 ? two / (float)(Ho - 1) : 0.f;
    P.step_d = Do > 1 ? two / (float)(Do - 1) : 0.f;
    P.rcp_w = one / (float)Wo; P.rcp_h = one / (float)Ho; P.rcp_d = one / (float)Do;
    P.exact_div = (Wo > DIVC_MAX_N || Ho > DIVC_MAX_N || Do > DIVC_MAX_N) ? 1 : 0;
    return P;
}

}  // namespace dgtta
