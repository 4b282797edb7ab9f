// GIN augmentation — tuned sm_100a path for the reference's configuration (gin_aug, dg_tta/gin.py:233-241:
// IN_CHANNELS=1, N_LAYER=4, INTERM_CHANNELS=2, kernel sizes drawn from {1,3}).
//
// The 4-layer stack (gin.py:94-113, 139-164) is cut into SEGMENTS, one launch per 3x3x3 layer:
//     [pointwise prologue: preceding 1x1x1 layers]  ->  3x3x3 conv  ->  [pointwise epilogue: following 1x1x1
//     layers, and on the last segment the alpha blend (gin.py:197) + sum-of-squares partials (gin.py:200-216)]
// so a 1x1x1 layer never costs an HBM round trip, and a stack without any 3x3x3 layer is one elementwise pass.
// Weights live in kernel parameters (constant bank): every FFMA takes its weight as a c[][] / uniform operand,
// no register or shared-memory traffic for them.  One launch per sample (weights differ per sample, groups=B).
//
// Conv kernel: a CTA owns a 32x32 (H x W) patch and marches along D.  Each new input plane is staged once in
// shared memory (zero padding of gin.py:105-107 folded into the staging; the prologue layers are applied there,
// outside-volume cells stay exactly 0).  A thread owns 4 consecutive outputs along W for all output channels and
// keeps three sets of accumulators (output planes d-1, d, d+1): an input plane is read once (3 rows x 6 words per
// input channel) and scattered into the three planes with packed FFMA2 (two adjacent outputs per instruction).
#include "common.cuh"

namespace dgtta {
namespace ginf {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int TH = 32, TW = 32, SW = 4;          // patch, outputs per thread along W
constexpr int NTHR = TH * (TW / SW);             // 256
constexpr int PH = TH + 2, PWD = 38;             // staged plane: 34 rows; pitch 38 words: rows alternate between banks
                                                 // {0,1 mod 4} and {2,3 mod 4} -> conflict-free LDS.64 per half-warp
constexpr int COL0 = 2;                          // smem column of w = w0-1 (even: every thread's 6-word window starts 8-byte aligned)
constexpr int MAXPW = 3;                         // pointwise layers in a prologue / epilogue chain

struct Pointwise {       // y_o = act(sum_i w[o][i] x_i + shift[o]); channels padded to 2 with zero weights
    float w[2][2];
    float shift[2];
    int act;
};

struct SegParams {
    const float *in;     // [rc, D, H, W] of this sample
    float *out;          // [oc, D, H, W] of this sample
    const float *x0;     // last segment: original input of this sample [1, D, H, W]
    double *partials;    // last segment: [RED_SLOTS][2] of this sample
    const float *alpha_ptr;
    int D, H, W;
    int nTH, nTW, nCD, chunkD;
    int rc;              // channels of `in` (1 or 2)
    int oc;              // channels of `out` (1 or 2)
    int n_pro, n_epi, last;
    int conv_act;
    Pointwise pro[MAXPW], epi[MAXPW];
    float cw[2 * 2 * 27];   // conv weights [cout][cin][kd][kh][kw] (reference layout of ker, gin.py:94)
    float cshift[2];
};

constexpr int RED_SLOTS = 1024;

__device__ __forceinline__ void apply_pointwise(const Pointwise &L, float &c0, float &c1)
{
    float y0 = fmaf(L.w[0][1], c1, L.w[0][0] * c0) + L.shift[0];
    float y1 = fmaf(L.w[1][1], c1, L.w[1][0] * c0) + L.shift[1];
    if (L.act) { y0 = y0 > 0.f ? y0 : y0 * 0.01f; y1 = y1 > 0.f ? y1 : y1 * 0.01f; }
    c0 = y0; c1 = y1;
}

__device__ __forceinline__ u64 lds64(const float *p) { return *reinterpret_cast<const u64 *>(p); }

// Order of the reference's conv accumulation is unspecified (cuDNN / mkldnn); here: cin, kh, kw ascending per input
// plane, planes scattered in the order kd = 2,1,0.
// NEPI / LAST are compile-time so that the per-output epilogue carries no runtime guards.
template <int CIN, int COUT, int NEPI, bool LAST>
__global__ void __launch_bounds__(NTHR) gin_conv_seg_kernel(const __grid_constant__ SegParams P)
{
    // Every staged plane is kept twice: tileA as is, tileB shifted left by one word.  The five two-output operand
    // pairs of a 3-tap window (x0x1, x1x2, x2x3, x3x4, x4x5) are then all 8-byte-aligned LDS.64 loads — no register
    // shuffling to build the odd pairs.
    __shared__ __align__(16) float tileA[2][CIN][PH * PWD];
    __shared__ __align__(16) float tileB[2][CIN][PH * PWD];
    __shared__ double red[2][NTHR / 32];
    const int tid = threadIdx.x;
    const int D = P.D, H = P.H, W = P.W;
    const size_t HW = (size_t)H * W, V = (size_t)D * HW;
    int bid = blockIdx.x;
    const int cd = bid % P.nCD; bid /= P.nCD;
    const int tw = bid % P.nTW; bid /= P.nTW;
    const int th = bid;
    const int h0 = th * TH, w0 = tw * TW;
    const int d0 = cd * P.chunkD, d1 = min(D, d0 + P.chunkD);

    // thread's outputs: row ty, columns 4*tx .. 4*tx+3
    const int ty = tid >> 3, tx = tid & 7;
    const int oh = h0 + ty, ow = w0 + SW * tx;
    const bool row_ok = oh < H;
    bool ok[SW];
#pragma unroll
    for (int k = 0; k < SW; ++k) ok[k] = row_ok && (ow + k < W);
    const bool vec_ok = ok[SW - 1] && ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);

    // staging: PH x (TW+2) cells per plane; cell i -> (row, col)
    constexpr int CELLS = PH * (TW + 2);
    constexpr int NC = (CELLS + NTHR - 1) / NTHR;
    int cs[NC], cg[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int i = tid + k * NTHR;
        cs[k] = -1; cg[k] = -1;
        if (i < CELLS) {
            const int rr = i / (TW + 2), cc = i - rr * (TW + 2);
            const int gh = h0 - 1 + rr, gw = w0 - 1 + cc;
            cs[k] = rr * PWD + COL0 + cc;
            cg[k] = (gh >= 0 && gh < H && gw >= 0 && gw < W) ? gh * W + gw : -1;   // -1: zero padding
        }
    }
    // staging is split so that the global loads of plane p+1 are in flight while plane p is being consumed
    float r0[NC], r1[NC];
    auto fetch = [&](int p) {   // plane p (inside the volume) -> registers
        const float *src = P.in + (size_t)p * HW;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            r0[k] = 0.f; r1[k] = 0.f;
            if (cg[k] >= 0) {
                r0[k] = __ldg(src + cg[k]);
                if (P.rc > 1) r1[k] = __ldg(src + V + cg[k]);
            }
        }
    };
    auto commit = [&](int buf) {   // registers -> tiles, applying the prologue layers inside the volume only
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (cs[k] < 0) continue;
            float c0 = r0[k], c1 = r1[k];
            if (cg[k] >= 0) {
#pragma unroll
                for (int l = 0; l < MAXPW; ++l)
                    if (l < P.n_pro) apply_pointwise(P.pro[l], c0, c1);
            }
            tileA[buf][0][cs[k]] = c0;
            tileB[buf][0][cs[k] - 1] = c0;
            if (CIN > 1) { tileA[buf][CIN - 1][cs[k]] = c1; tileB[buf][CIN - 1][cs[k] - 1] = c1; }
        }
    };

    u64 acc[3][COUT][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[s][o][0] = acc[s][o][1] = 0ull;
    double s_in = 0.0, s_mix = 0.0;
    float alpha = 0.f;
    if (LAST) alpha = __ldg(P.alpha_ptr);

    // input planes p = d0-1 .. d1 ; plane p completes output plane p-1
    const int p_begin = d0 - 1, p_end = d1 + 1;
    if (p_begin >= 0) { fetch(p_begin); commit(0); }
    __syncthreads();
    const int woff = ty * PWD + COL0 + SW * tx;   // row ty-1+kh (tile row ty+kh), column of w-1: even
    for (int pb = p_begin; pb < p_end; pb += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int p = pb + u;
            if (p >= p_end) break;
            const int buf = (p - p_begin) & 1;
            const bool have_next = p + 1 < p_end && p + 1 < D;
            if (have_next) fetch(p + 1);
            // slots: output plane q uses slot (q - p_begin + 3) % 3; with p = p_begin + 3m + u:
            //   q = p+1 -> slot (u+1)%3 (tap kd=0) , q = p -> slot u (kd=1) , q = p-1 -> slot (u+2)%3 (kd=2)
            if (p >= 0 && p < D) {
#pragma unroll
                for (int ic = 0; ic < CIN; ++ic) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float *ra = &tileA[buf][ic][woff + kh * PWD];
                        const float *rb = &tileB[buf][ic][woff + kh * PWD];
                        const u64 P01 = lds64(ra), P23 = lds64(ra + 2), P45 = lds64(ra + 4);
                        const u64 P12 = lds64(rb), P34 = lds64(rb + 2);
#pragma unroll
                        for (int kd = 0; kd < 3; ++kd) {
                            const int slot = (kd == 0) ? (u + 1) % 3 : (kd == 1 ? u : (u + 2) % 3);
#pragma unroll
                            for (int o = 0; o < COUT; ++o) {
                                const float *wv = &P.cw[((o * CIN + ic) * 3 + kd) * 9 + kh * 3];
                                u64 a0 = acc[slot][o][0], a1 = acc[slot][o][1];
                                a0 = ffma2(P01, pk(wv[0], wv[0]), a0); a1 = ffma2(P23, pk(wv[0], wv[0]), a1);
                                a0 = ffma2(P12, pk(wv[1], wv[1]), a0); a1 = ffma2(P34, pk(wv[1], wv[1]), a1);
                                a0 = ffma2(P23, pk(wv[2], wv[2]), a0); a1 = ffma2(P45, pk(wv[2], wv[2]), a1);
                                acc[slot][o][0] = a0; acc[slot][o][1] = a1;
                            }
                        }
                    }
                }
            }
            // output plane q = p-1 is complete (its three input planes p-2, p-1, p have been scattered)
            const int q = p - 1;
            const int qs = (u + 2) % 3;
            if (q >= d0 && q < d1) {
                float y[2][SW];
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    unpk(acc[qs][o][0], y[o][0], y[o][1]);
                    unpk(acc[qs][o][1], y[o][2], y[o][3]);
                }
                if (COUT == 1) {
#pragma unroll
                    for (int k = 0; k < SW; ++k) y[1][k] = 0.f;
                }
                const size_t off = (size_t)q * HW + (size_t)oh * W + ow;
                float xin[SW];
                if (LAST) {
                    if (vec_ok) {
                        const float4 xv = __ldg(reinterpret_cast<const float4 *>(P.x0 + off));
                        xin[0] = xv.x; xin[1] = xv.y; xin[2] = xv.z; xin[3] = xv.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < SW; ++k) xin[k] = ok[k] ? __ldg(P.x0 + off + k) : 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < SW; ++k) {
                    float c0 = y[0][k] + P.cshift[0], c1 = y[1][k] + P.cshift[1];        // gin.py:111
                    if (!(LAST && NEPI == 0)) {   // the conv layer itself is the stack's last layer only then: no activation
                        c0 = c0 > 0.f ? c0 : c0 * 0.01f; c1 = c1 > 0.f ? c1 : c1 * 0.01f;   // gin.py:112-113
                    }
                    if (COUT == 1) c1 = 0.f;
#pragma unroll
                    for (int l = 0; l < NEPI; ++l) apply_pointwise(P.epi[l], c0, c1);
                    if (LAST) {
                        c0 = __fadd_rn(__fmul_rn(alpha, c0), __fmul_rn(1.0f - alpha, xin[k]));   // gin.py:197
                        if (ok[k]) { s_in += (double)xin[k] * xin[k]; s_mix += (double)c0 * c0; }
                    }
                    y[0][k] = c0; y[1][k] = c1;
                }
                if (vec_ok) {
                    *reinterpret_cast<float4 *>(P.out + off) = make_float4(y[0][0], y[0][1], y[0][2], y[0][3]);
                    if (!LAST && P.oc > 1) *reinterpret_cast<float4 *>(P.out + V + off) = make_float4(y[1][0], y[1][1], y[1][2], y[1][3]);
                } else {
#pragma unroll
                    for (int k = 0; k < SW; ++k)
                        if (ok[k]) {
                            P.out[off + k] = y[0][k];
                            if (!LAST && P.oc > 1) P.out[V + off + k] = y[1][k];
                        }
                }
            }
            // recycle the finished slot for output plane p+2
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[qs][o][0] = acc[qs][o][1] = 0ull;
            if (have_next) commit(buf ^ 1);   // the other buffer was last read one plane ago (barrier below, previous turn)
            __syncthreads();   // next plane staged by everyone; this plane's buffer free for re-staging
        }
    }
    if (LAST) {
        s_in = warp_sum(s_in); s_mix = warp_sum(s_mix);
        if ((tid & 31) == 0) { red[0][tid >> 5] = s_in; red[1][tid >> 5] = s_mix; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, m = 0.0;
            for (int i = 0; i < NTHR / 32; ++i) { a += red[0][i]; m += red[1][i]; }
            const int slot = blockIdx.x % RED_SLOTS;
            atomicAdd(&P.partials[2 * slot], a);
            atomicAdd(&P.partials[2 * slot + 1], m);
        }
    }
}

template <int CIN, int COUT>
static void launch_conv(const SegParams &P, unsigned grid, cudaStream_t stream)
{
    // the reachable (NEPI, LAST) combinations: the epilogue holds the 1x1x1 layers up to the next 3x3x3 layer or the end
    const int code = P.n_epi * 2 + (P.last ? 1 : 0);
    switch (code) {
        case 0: gin_conv_seg_kernel<CIN, COUT, 0, false><<<grid, NTHR, 0, stream>>>(P); break;
        case 1: gin_conv_seg_kernel<CIN, COUT, 0, true><<<grid, NTHR, 0, stream>>>(P); break;
        case 2: gin_conv_seg_kernel<CIN, COUT, 1, false><<<grid, NTHR, 0, stream>>>(P); break;
        case 3: gin_conv_seg_kernel<CIN, COUT, 1, true><<<grid, NTHR, 0, stream>>>(P); break;
        case 4: gin_conv_seg_kernel<CIN, COUT, 2, false><<<grid, NTHR, 0, stream>>>(P); break;
        case 5: gin_conv_seg_kernel<CIN, COUT, 2, true><<<grid, NTHR, 0, stream>>>(P); break;
        default: gin_conv_seg_kernel<CIN, COUT, 3, true><<<grid, NTHR, 0, stream>>>(P); break;
    }
}

// stack without any 3x3x3 layer: one elementwise pass (4 voxels per thread)
__global__ void __launch_bounds__(256) gin_pointwise_kernel(const __grid_constant__ SegParams P)
{
    __shared__ double red[2][8];
    const size_t V = (size_t)P.D * P.H * P.W;
    const float alpha = __ldg(P.alpha_ptr);
    double s_in = 0.0, s_mix = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += stride) {
        const float x = __ldg(P.in + i);
        float c0 = x, c1 = 0.f;
#pragma unroll
        for (int l = 0; l < MAXPW + 1; ++l)
            if (l < P.n_epi) apply_pointwise(l < MAXPW ? P.epi[l] : P.pro[0], c0, c1);
        c0 = __fadd_rn(__fmul_rn(alpha, c0), __fmul_rn(1.0f - alpha, x));   // gin.py:197
        s_in += (double)x * x; s_mix += (double)c0 * c0;
        P.out[i] = c0;
    }
    s_in = warp_sum(s_in); s_mix = warp_sum(s_mix);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_in; red[1][threadIdx.x >> 5] = s_mix; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, m = 0.0;
        for (int i = 0; i < 8; ++i) { a += red[0][i]; m += red[1][i]; }
        const int slot = blockIdx.x % RED_SLOTS;
        atomicAdd(&P.partials[2 * slot], a);
        atomicAdd(&P.partials[2 * slot + 1], m);
    }
}

template <int CIN, int COUT>
static void touch_conv()
{
    DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 0, false>); DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 0, true>);
    DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 1, false>); DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 1, true>);
    DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 2, false>); DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 2, true>);
    DGTTA_TOUCH(gin_conv_seg_kernel<CIN, COUT, 3, true>);
}

static void fill_pointwise(Pointwise &L, const float *ker, const float *shift, int cin, int cout, int act)
{
    for (int o = 0; o < 2; ++o) {
        for (int i = 0; i < 2; ++i) L.w[o][i] = (o < cout && i < cin) ? ker[o * cin + i] : 0.f;
        L.shift[o] = o < cout ? shift[o] : 0.f;
    }
    L.act = act;
}

}  // namespace ginf

void preload_gin_fused()
{
    ginf::touch_conv<1, 2>(); ginf::touch_conv<2, 2>(); ginf::touch_conv<2, 1>();
    DGTTA_TOUCH(ginf::gin_pointwise_kernel);
}

// Runs the tuned path for cfg (1, 4, 2).  params_host layout as in dgtta_gin_fwd.  bufs: two device buffers of
// B*2*V floats.  partials: [B][RED_SLOTS][2] doubles, zeroed by the caller.
int gin_fused_launch(const float *x_dev, float *out_dev, const float *params_host, const int *ks, const float *alphas_dev,
                     int B, int D, int H, int W, float *buf0, float *buf1, double *partials, cudaStream_t stream)
{
    using namespace ginf;
    const int cins[4] = {1, 2, 2, 2}, couts[4] = {2, 2, 2, 1};
    const size_t V = (size_t)D * H * W;
    // per-layer offsets into params_host
    size_t koff[4], soff[4], off = 0;
    for (int L = 0; L < 4; ++L) {
        const int k3 = ks[L] * ks[L] * ks[L];
        koff[L] = off; off += (size_t)couts[L] * B * cins[L] * k3;
        soff[L] = off; off += (size_t)couts[L] * B;
    }
    int convs[4], nconv = 0;
    for (int L = 0; L < 4; ++L) if (ks[L] == 3) convs[nconv++] = L;

    SegParams P;
    P.D = D; P.H = H; P.W = W;
    P.nTH = (H + TH - 1) / TH; P.nTW = (W + TW - 1) / TW;
    // D chunks: aim at >= 4 CTAs per SM while keeping the 2 halo planes a small fraction of a chunk
    const long base = (long)P.nTH * P.nTW;
    int ncd = (int)((4L * sm_count() + base - 1) / base);
    const int max_chunks = (D + 11) / 12;
    if (ncd > max_chunks) ncd = max_chunks;
    if (ncd < 1) ncd = 1;
    P.chunkD = (D + ncd - 1) / ncd;
    P.nCD = (D + P.chunkD - 1) / P.chunkD;

    for (int b = 0; b < B; ++b) {
        P.alpha_ptr = alphas_dev + b;
        P.partials = partials + (size_t)b * RED_SLOTS * 2;
        P.x0 = x_dev + (size_t)b * V;
        auto ker = [&](int L) { return params_host + koff[L] + (size_t)b * couts[L] * cins[L] * ks[L] * ks[L] * ks[L]; };
        auto shf = [&](int L) { return params_host + soff[L] + (size_t)b * couts[L]; };
        if (nconv == 0) {
            // epi[0..2] = layers 0..2, pro[0] = layer 3 (the chain has 4 pointwise layers)
            for (int L = 0; L < 3; ++L) fill_pointwise(P.epi[L], ker(L), shf(L), cins[L], couts[L], 1);
            fill_pointwise(P.pro[0], ker(3), shf(3), cins[3], couts[3], 0);
            P.n_epi = 4; P.n_pro = 0; P.last = 1; P.rc = 1; P.oc = 1; P.conv_act = 0;
            P.in = x_dev + (size_t)b * V; P.out = out_dev + (size_t)b * V;
            size_t gx = (V + 255) / 256;
            const size_t cap = (size_t)sm_count() * 16;
            if (gx > cap) gx = cap;
            gin_pointwise_kernel<<<(unsigned)gx, 256, 0, stream>>>(P);
            int rc = check_launch("gin_pointwise_kernel");
            if (rc) return rc;
            continue;
        }
        const float *cur = x_dev + (size_t)b * V;
        int cur_c = 1;
        for (int s = 0; s < nconv; ++s) {
            const int Lc = convs[s];
            const int first = (s == 0) ? 0 : convs[s - 1] + 1;   // pointwise layers before the conv that are still pending
            // layers between the previous conv and this one were folded into the previous segment's epilogue,
            // except for segment 0 whose leading pointwise layers become the prologue
            P.n_pro = 0;
            if (s == 0)
                for (int L = first; L < Lc; ++L) fill_pointwise(P.pro[P.n_pro++], ker(L), shf(L), cins[L], couts[L], 1);
            const int last_seg = (s == nconv - 1);
            const int epi_end = last_seg ? 4 : convs[s + 1];
            P.n_epi = 0;
            for (int L = Lc + 1; L < epi_end; ++L) fill_pointwise(P.epi[P.n_epi++], ker(L), shf(L), cins[L], couts[L], L != 3);
            for (int i = 0; i < 2 * 2 * 27; ++i) P.cw[i] = 0.f;
            const int ci = cins[Lc], co = couts[Lc];
            for (int i = 0; i < co * ci * 27; ++i) P.cw[i] = ker(Lc)[i];
            P.cshift[0] = shf(Lc)[0]; P.cshift[1] = co > 1 ? shf(Lc)[1] : 0.f;
            P.conv_act = (Lc != 3);
            P.last = last_seg;
            P.rc = cur_c;
            P.oc = last_seg ? 1 : couts[epi_end - 1];
            P.in = cur;
            float *dst = last_seg ? out_dev + (size_t)b * V : ((s & 1) ? buf1 : buf0) + (size_t)b * 2 * V;
            P.out = dst;
            const unsigned grid = (unsigned)(base * P.nCD);
            if (ci == 1 && co == 2) launch_conv<1, 2>(P, grid, stream);
            else if (ci == 2 && co == 2) launch_conv<2, 2>(P, grid, stream);
            else launch_conv<2, 1>(P, grid, stream);
            int rc = check_launch("gin_conv_seg_kernel");
            if (rc) return rc;
            cur = dst;
            cur_c = P.oc;
        }
    }
    return 0;
}

}  // namespace dgtta
