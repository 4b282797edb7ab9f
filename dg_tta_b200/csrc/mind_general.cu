// MIND-SSC descriptor - general path (any tap radius 1..4, any dilation).  See mind_ssc.cu for the math and
// mind_fast.cu for the tuned kernel that serves the reference's default (5 taps, delta <= 3).
//
// Kernel structure ("plane marching"): one CTA owns a TH x TW patch of the H-W plane and walks along D.
// For every plane it (S1) evaluates the 12 squared edge channels on the patch plus a halo of R into
// shared memory, (S2) smooths them along W with register-blocked runs, (S3) smooths along H while
// reading the thread's own column, and keeps the last 2R+1 in-plane results of its voxel in registers
// so that the smoothing along D never touches memory.  Pass 1 assumes the global clamp of
// mind.py:158-160 inactive and records per CTA {sum v, min positive v, max v}; pass 2 (FIX) reduces
// those to mean_all(v) and recomputes only CTAs in which some v leaves [0.001*mean, 1000*mean].
#include "mind_internal.cuh"
#include <atomic>


namespace dgtta {

constexpr int MIND_RUN = 8;                      // W-pass outputs per task

template <int R>
struct MindGeom {
    static constexpr int NT = 2 * R + 1;
    static constexpr int EH = MIND_TH + 2 * R;
    static constexpr int EW = MIND_TW + 2 * R;
    static constexpr int QUADS = (MIND_RUN + 2 * R + 3) / 4;            // float4 loads per W-pass task
    static constexpr int EWP = (MIND_TW - MIND_RUN) + 4 * QUADS;        // padded row so the last task stays in-row
    static constexpr int NPOS = (EH * EW + MIND_THREADS - 1) / MIND_THREADS;
    static constexpr int NTASK = 12 * EH * (MIND_TW / MIND_RUN);
    static constexpr int E2_FLOATS = 12 * EH * EWP;
    static constexpr int WS_FLOATS = 12 * EH * MIND_TW;
    static constexpr size_t SMEM = sizeof(float) * (E2_FLOATS + WS_FLOATS);
};

struct MindParams {
    const float *img;
    float *out;
    const float *noise;
    const float *in_scale;
    float4 *tile_stats;  // per CTA: {sum v, min positive v, max v, -}
    int B, D, H, W;
    int delta;
    int nTH, nTW, nCD, chunkD;
    int noise_mode;
    float rw;
    float taps[9];
    double inv_count;  // 1 / (B*D*H*W)
};

template <int R, bool FIX, int NOISE>
__global__ void __launch_bounds__(MIND_THREADS, 1) mind_general_kernel(const __grid_constant__ MindParams P)
{
    using G = MindGeom<R>;
    constexpr int NT = G::NT;
    extern __shared__ __align__(16) float smem[];
    float *e2 = smem;                  // [12][EH][EWP]  squared edges of the current plane
    float *ws = smem + G::E2_FLOATS;   // [12][EH][TW]   after the W pass
    __shared__ double red_d[MIND_THREADS / 32];
    __shared__ float red_f[3][MIND_THREADS / 32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    int bid = blockIdx.x;
    const int ntiles = gridDim.x;
    const int cd = bid % P.nCD; bid /= P.nCD;
    const int tw = bid % P.nTW; bid /= P.nTW;
    const int th = bid % P.nTH; bid /= P.nTH;
    const int b = bid;
    const int H = P.H, W = P.W, D = P.D;
    const int HW = H * W;

    float lo = 0.f, hi = 0.f;
    if (FIX) {
        // mean_all(v) from the per-tile sums, same order in every CTA -> identical value everywhere
        double s = 0.0;
        for (int i = tid; i < ntiles; i += MIND_THREADS) s += (double)P.tile_stats[i].x;
        s = warp_sum(s);
        if (lane == 0) red_d[warp] = s;
        __syncthreads();
        double tot = 0.0;
        for (int i = 0; i < MIND_THREADS / 32; ++i) tot += red_d[i];
        const float mean = (float)(tot * P.inv_count);
        lo = mean * 0.001f;  // mind.py:158-160
        hi = mean * 1000.f;
        const float4 st = P.tile_stats[blockIdx.x];
        const bool need = !(mean > 0.f) || st.z > hi || st.y < lo;
        if (!need) return;
    }

    const int h0 = th * MIND_TH, w0 = tw * MIND_TW;
    const int d0 = cd * P.chunkD;
    const int d1 = min(D, d0 + P.chunkD);
    const float *img = P.img + (size_t)b * D * HW;
    float sa = 1.f, sc = 1.f;
    const bool scaled = P.in_scale != nullptr;
    if (scaled) { sa = P.in_scale[2 * b]; sc = P.in_scale[2 * b + 1]; }
    const int delta = P.delta;

    // S1 bookkeeping: each thread owns up to NPOS halo positions; offsets are plane-invariant
    int s1_smem[G::NPOS], o_c[G::NPOS], o_hm[G::NPOS], o_hp[G::NPOS], o_wm[G::NPOS], o_wp[G::NPOS];
#pragma unroll
    for (int k = 0; k < G::NPOS; ++k) {
        const int i = tid + k * MIND_THREADS;
        if (i < G::EH * G::EW) {
            const int row = i / G::EW, col = i - row * G::EW;
            // E^2 outside the volume is E^2 at the clamped position (replicate padding of the smoothing,
            // mind.py:22), so clamp the centre first and then apply the (clamped) shifts
            const int gh = clampi(h0 - R + row, 0, H - 1), gw = clampi(w0 - R + col, 0, W - 1);
            s1_smem[k] = row * G::EWP + col;
            o_c[k] = gh * W + gw;
            o_hm[k] = clampi(gh - delta, 0, H - 1) * W + gw;
            o_hp[k] = clampi(gh + delta, 0, H - 1) * W + gw;
            o_wm[k] = gh * W + clampi(gw - delta, 0, W - 1);
            o_wp[k] = gh * W + clampi(gw + delta, 0, W - 1);
        } else {
            s1_smem[k] = -1;
            o_c[k] = o_hm[k] = o_hp[k] = o_wm[k] = o_wp[k] = 0;
        }
    }

    float taps[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) taps[t] = P.taps[t];

    const int ty = tid / MIND_TW, tx = tid - ty * MIND_TW;
    const int vh = h0 + ty, vw = w0 + tx;
    const bool valid = vh < H && vw < W;
    float *outp = P.out + ((size_t)b * 12 * D) * HW + (size_t)vh * W + vw;

    float win[12][NT];
#pragma unroll
    for (int c = 0; c < 12; ++c)
#pragma unroll
        for (int t = 0; t < NT; ++t) win[c][t] = 0.f;
    float x[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) x[c] = 0.f;

    float st_sum = 0.f, st_min = __int_as_float(0x7f800000), st_max = 0.f;

    int prev_zc = -1;
    const int z_begin = d0 - R, z_end = d1 + R;  // logical planes pushed into the D window
    for (int zb = z_begin; zb < z_end; zb += NT) {
#pragma unroll
        for (int ph = 0; ph < NT; ++ph) {
            const int z = zb + ph;
            if (z < z_end) {
                const int zc = clampi(z, 0, D - 1);
                if (zc != prev_zc) {  // uniform over the CTA
                    prev_zc = zc;
                    // ---------------- S1: squared edges of plane zc on the halo patch
                    const float *p0 = img + (size_t)zc * HW;
                    const float *pm = img + (size_t)clampi(zc - delta, 0, D - 1) * HW;
                    const float *pp = img + (size_t)clampi(zc + delta, 0, D - 1) * HW;
#pragma unroll
                    for (int k = 0; k < G::NPOS; ++k) {
                        if (s1_smem[k] >= 0) {
                            float nb[6];
                            nb[NB_DM] = __ldg(pm + o_c[k]);
                            nb[NB_DP] = __ldg(pp + o_c[k]);
                            nb[NB_HM] = __ldg(p0 + o_hm[k]);
                            nb[NB_HP] = __ldg(p0 + o_hp[k]);
                            nb[NB_WM] = __ldg(p0 + o_wm[k]);
                            nb[NB_WP] = __ldg(p0 + o_wp[k]);
                            if (scaled) {
#pragma unroll
                                for (int q = 0; q < 6; ++q) nb[q] = __fmul_rn(__fmul_rn(nb[q], sa), sc);
                            }
                            float *dst = e2 + s1_smem[k];
#pragma unroll
                            for (int c = 0; c < 12; ++c) {
                                float e = nb[mind_p1(c)] - nb[mind_p2(c)];
                                if (NOISE == DGTTA_NOISE_TENSOR) {
                                    const float n = __ldg(P.noise + ((size_t)(b * 12 + c) * D + zc) * HW + o_c[k]);
                                    e = __fadd_rn(e, __fmul_rn(P.rw, n));  // mind.py:150-152, two roundings
                                }
                                dst[c * (G::EH * G::EWP)] = e * e;
                            }
                        }
                    }
                    __syncthreads();
                    // ---------------- S2: smooth along W, MIND_RUN outputs per task
                    for (int t = tid; t < G::NTASK; t += MIND_THREADS) {
                        const int j = t & (MIND_TW / MIND_RUN - 1);
                        const int cr = t / (MIND_TW / MIND_RUN);  // c * EH + row
                        const float4 *src = reinterpret_cast<const float4 *>(e2 + cr * G::EWP + j * MIND_RUN);
                        float v[4 * G::QUADS];
#pragma unroll
                        for (int q = 0; q < G::QUADS; ++q) {
                            const float4 f = src[q];
                            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
                        }
                        float o[MIND_RUN];
#pragma unroll
                        for (int k = 0; k < MIND_RUN; ++k) {
                            float acc = taps[0] * v[k];
#pragma unroll
                            for (int tt = 1; tt < NT; ++tt) acc = fmaf(taps[tt], v[k + tt], acc);
                            o[k] = acc;
                        }
                        float4 *dst = reinterpret_cast<float4 *>(ws + cr * MIND_TW + j * MIND_RUN);
                        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                    }
                    __syncthreads();
                    // ---------------- S3: smooth along H for the thread's own voxel
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        const float *col = ws + (c * G::EH + ty) * MIND_TW + tx;
                        float acc = taps[0] * col[0];
#pragma unroll
                        for (int tt = 1; tt < NT; ++tt) acc = fmaf(taps[tt], col[tt * MIND_TW], acc);
                        x[c] = acc;
                    }
                    // the next S1 overwrites e2 only after every thread passed the S2 barrier of this plane,
                    // and the next S2 overwrites ws only after the next S1 barrier, which every thread reaches
                    // after finishing this S3: no extra barrier needed.
                }
                // ---------------- D pass: push x into the register window (slot ph), emit plane z-R
#pragma unroll
                for (int c = 0; c < 12; ++c) win[c][ph] = x[c];
                const int d = z - R;
                if (d >= d0) {
                    float ssd[12];
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        float acc = taps[0] * win[c][(ph + 1) % NT];
#pragma unroll
                        for (int tt = 1; tt < NT; ++tt) acc = fmaf(taps[tt], win[c][(ph + 1 + tt) % NT], acc);
                        ssd[c] = acc;
                    }
                    float mn = ssd[0];
#pragma unroll
                    for (int c = 1; c < 12; ++c) mn = fminf(mn, ssd[c]);
                    float sum = 0.f;
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        ssd[c] -= mn;   // mind.py:156
                        sum += ssd[c];
                    }
                    float v = __fdiv_rn(sum, 12.f);  // mind.py:157
                    float scale;
                    if (FIX) {
                        v = fminf(fmaxf(v, lo), hi);  // mind.py:158-160
                        scale = -1.4426950408889634f * __fdiv_rn(1.f, v);
                    } else {
                        if (valid) {
                            st_sum += v;
                            st_max = fmaxf(st_max, v);
                            if (v > 0.f) st_min = fminf(st_min, v);
                        }
                        // v == 0 means every m_c == 0: exp(-0/lo) = 1 for any lo > 0 (lo == 0 is caught by pass 2)
                        scale = v > 0.f ? -1.4426950408889634f * __fdiv_rn(1.f, v) : 0.f;
                    }
                    if (valid) {
                        float *o = outp + (size_t)d * HW;
#pragma unroll
                        for (int c = 0; c < 12; ++c)
                            __stcs(o + (size_t)c * D * HW, ex2_approx(ssd[c] * scale));  // mind.py:161-162
                    }
                }
            }
        }
    }

    if (!FIX) {
        st_sum = warp_sum(st_sum);
        st_min = warp_min(st_min);
        st_max = warp_max(st_max);
        if (lane == 0) { red_f[0][warp] = st_sum; red_f[1][warp] = st_min; red_f[2][warp] = st_max; }
        __syncthreads();
        if (tid == 0) {
            float s = 0.f, mnv = __int_as_float(0x7f800000), mxv = 0.f;
            for (int i = 0; i < MIND_THREADS / 32; ++i) {
                s += red_f[0][i];
                mnv = fminf(mnv, red_f[1][i]);
                mxv = fmaxf(mxv, red_f[2][i]);
            }
            P.tile_stats[blockIdx.x] = make_float4(s, mnv, mxv, 0.f);
        }
    }
}

struct MindPlan {
    int nTH, nTW, nCD, chunkD, ntiles;
};

static MindPlan mind_plan(int B, int D, int H, int W)
{
    MindPlan p;
    p.nTH = (H + MIND_TH - 1) / MIND_TH;
    p.nTW = (W + MIND_TW - 1) / MIND_TW;
    const long base = (long)B * p.nTH * p.nTW;
    // one CTA per SM is resident (512 threads, ~65 KB smem): split D so that the grid fills the SMs,
    // but keep chunks long enough that the 2R warm-up planes stay a small fraction of the work
    const int sms = sm_count();
    int ncd = (int)((sms + base - 1) / base);
    const int max_chunks = (D + 15) / 16;
    if (ncd > max_chunks) ncd = max_chunks;
    if (ncd < 1) ncd = 1;
    p.chunkD = (D + ncd - 1) / ncd;
    p.nCD = (D + p.chunkD - 1) / p.chunkD;
    p.ntiles = (int)(base * p.nCD);
    return p;
}

template <int R, int NOISE>
static int mind_launch(const MindParams &P, int ntiles, cudaStream_t stream)
{
    using G = MindGeom<R>;
    // the opt-in to > 48 KB of dynamic shared memory is per device: remember it per (instantiation, device)
    static std::atomic<bool> configured_on[64];   // idempotent set-up: a race only repeats it
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::atomic<bool> &configured = configured_on[dev];
    if (!configured) {
        cudaFuncSetAttribute(mind_general_kernel<R, false, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        cudaFuncSetAttribute(mind_general_kernel<R, true, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        configured = true;
    }
    mind_general_kernel<R, false, NOISE><<<ntiles, MIND_THREADS, G::SMEM, stream>>>(P);
    int rc = check_launch("mind_general_kernel<spec>");
    if (rc) return rc;
    mind_general_kernel<R, true, NOISE><<<ntiles, MIND_THREADS, G::SMEM, stream>>>(P);
    return check_launch("mind_general_kernel<fix>");
}

template <int R>
static int mind_dispatch_noise(const MindParams &P, int ntiles, cudaStream_t stream)
{
    switch (P.noise_mode) {
        case DGTTA_NOISE_NONE: return mind_launch<R, DGTTA_NOISE_NONE>(P, ntiles, stream);
        case DGTTA_NOISE_TENSOR: return mind_launch<R, DGTTA_NOISE_TENSOR>(P, ntiles, stream);
        default:
            set_error("dgtta_mind_ssc_fwd: noise_mode %d not supported", P.noise_mode);
            return DGTTA_EUNSUPPORTED;
    }
}


template <int R>
static void touch_general()
{
    DGTTA_TOUCH(mind_general_kernel<R, false, DGTTA_NOISE_NONE>); DGTTA_TOUCH(mind_general_kernel<R, true, DGTTA_NOISE_NONE>);
    DGTTA_TOUCH(mind_general_kernel<R, false, DGTTA_NOISE_TENSOR>); DGTTA_TOUCH(mind_general_kernel<R, true, DGTTA_NOISE_TENSOR>);
}

void preload_mind_general()
{
    // the reference's configuration (5 taps) runs in mind_fast.cu; of the general path only the 5-tap instance (large
    // delta) is preloaded, the 3/7/9-tap instances keep CUDA's lazy loading
    touch_general<2>();
}

size_t mind_general_workspace_bytes(int B, int D, int H, int W)
{
    // upper bound independent of the SM count: at most ceil(D/16) chunks (see mind_plan)
    const size_t tiles = (size_t)B * ((H + MIND_TH - 1) / MIND_TH) * ((W + MIND_TW - 1) / MIND_TW) * ((D + 15) / 16);
    return tiles * sizeof(float4);
}

int mind_general_launch(const MindArgs &a, cudaStream_t stream)
{
    const MindPlan plan = mind_plan(a.B, a.D, a.H, a.W);
    if (a.workspace_bytes < (size_t)plan.ntiles * sizeof(float4)) {
        set_error("dgtta_mind_ssc_fwd: workspace too small (%zu < %zu)", a.workspace_bytes, (size_t)plan.ntiles * sizeof(float4));
        return DGTTA_EWORKSPACE;
    }
    MindParams P;
    P.img = a.img; P.out = a.out; P.noise = a.noise; P.in_scale = a.in_scale;
    P.tile_stats = (float4 *)a.workspace;
    P.B = a.B; P.D = a.D; P.H = a.H; P.W = a.W; P.delta = a.delta;
    P.nTH = plan.nTH; P.nTW = plan.nTW; P.nCD = plan.nCD; P.chunkD = plan.chunkD;
    P.noise_mode = a.noise_mode; P.rw = a.rw;
    for (int i = 0; i < 9; ++i) P.taps[i] = a.taps[i];
    P.inv_count = 1.0 / ((double)a.B * a.D * a.H * a.W);
    switch (a.ntaps / 2) {
        case 1: return mind_dispatch_noise<1>(P, plan.ntiles, stream);
        case 2: return mind_dispatch_noise<2>(P, plan.ntiles, stream);
        case 3: return mind_dispatch_noise<3>(P, plan.ntiles, stream);
        case 4: return mind_dispatch_noise<4>(P, plan.ntiles, stream);
        default: set_error("dgtta_mind_ssc_fwd: %d taps not supported", a.ntaps); return DGTTA_EUNSUPPORTED;
    }
}

}  // namespace dgtta
