// Low-resolution simulation of the MultiRes trainers: the two resampling steps of
// dg_tta/pretraining/discrete_downsampling.py:8-37 (augment_discrete_linear_downsampling_scipy), used at
// nnUNetTrainer_GIN_MIND_MultiRes.py:57-69 with order_downsample = 0 (nearest) and order_upsample = 3 (cubic):
//     downsampled = resize(x, target_shape, order=order_downsample, mode='edge', anti_aliasing=False)
//     x           = resize(downsampled, shp, order=order_upsample,  mode='edge', anti_aliasing=False)
// skimage.transform.resize is third party (not vendored under /root/reference; scikit-image >= 0.19 via nnunetv2 2.2.1) and
// delegates to scipy.ndimage.zoom(image, out/in, order, mode='nearest', grid_mode=True), then clips to the input's range.
// The arithmetic restated here is scipy's (ni_interpolation.c NI_ZoomShift, ni_splines.c); the CPU checker of the tests is
// pinned in the build container against scipy.ndimage.zoom 1.18 itself (see the tests):
//   coordinate of output index o on an axis of n_in -> n_out samples (double, this operation order):
//       cc = ((o + 0.5) * (n_in / n_out)) - 0.5
//   order 0: index floor(clamp(cc, 0, n_in-1) + 0.5);  order 1: taps floor(cc), +1 with weights (1-y, y), taps clamped;
//   order 3: the input is edge-padded by 12 samples and pre-filtered into cubic B-spline coefficients (gain 6, pole
//       sqrt(3)-2, causal / anticausal recursion with the half-sample-symmetric initialisation scipy uses for mode
//       'nearest'), then  out = sum_{ijk} c[s_d+i][s_h+j][s_w+k] w_d[i] w_h[j] w_w[k],  s = floor(cc) - 1 + 12, cc NOT
//       clamped, w = cubic B-spline weights of y = cc - floor(cc);
//   result clipped to [min(input), max(input)] (skimage clip=True).
// The per-axis tables (index / start + weights) are evaluated on the host in double with exactly these operations and
// uploaded; spline coefficients and the 64-tap sums are double precision on the device (the volumes are small: the
// low-resolution image is at most 1/8 of the patch), the result is rounded to float32 once, like the reference's
// assignment into its float32 array.
#include <math.h>
#include <vector>

#include "common.cuh"

namespace dgtta {
namespace rsz {

constexpr int NPAD = 12;      // scipy _prepad_for_spline_filter

struct Geo {
    int N, Di, Hi, Wi, Do, Ho, Wo;
};

// min / max of every input volume (np.clip bounds); one block per volume
__global__ void __launch_bounds__(1024) minmax_kernel(const float *in, long long n, float *mm)
{
    __shared__ float smin[32], smax[32];
    const float *x = in + (size_t)blockIdx.x * n;
    float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
    for (long long i = threadIdx.x; i < n; i += 1024) { const float v = x[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = warp_min(smin[threadIdx.x]); hi = warp_max(smax[threadIdx.x]);
        if (threadIdx.x == 0) { mm[2 * blockIdx.x] = lo; mm[2 * blockIdx.x + 1] = hi; }
    }
}

// order 0 / 1: tables idx[axis][o] (first tap, already clamped for order 0) and, for order 1, wgt[axis][o] = y
template <int ORDER>
__global__ void __launch_bounds__(256) gather_kernel(const float *in, float *out, const int *idx, const double *wy, const float *mm, Geo G)
{
    const long long Vo = (long long)G.Do * G.Ho * G.Wo, Vi = (long long)G.Di * G.Hi * G.Wi;
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (v >= Vo) return;
    const int w = (int)(v % G.Wo), h = (int)((v / G.Wo) % G.Ho), d = (int)(v / ((long long)G.Wo * G.Ho));
    const float *x = in + (size_t)n * Vi;
    const int id = idx[d], ih = idx[G.Do + h], iw = idx[G.Do + G.Ho + w];
    double r;
    if (ORDER == 0) {
        r = x[((size_t)id * G.Hi + ih) * G.Wi + iw];
    } else {
        const double yd = wy[d], yh = wy[G.Do + h], yw = wy[G.Do + G.Ho + w];
        r = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int dd = min(max(id + i, 0), G.Di - 1), hh = min(max(ih + j, 0), G.Hi - 1), ww = min(max(iw + k, 0), G.Wi - 1);
                    const double wt = (i ? yd : 1.0 - yd) * (j ? yh : 1.0 - yh) * (k ? yw : 1.0 - yw);
                    r += (double)x[((size_t)dd * G.Hi + hh) * G.Wi + ww] * wt;
                }
    }
    const double lo = mm[2 * n], hi = mm[2 * n + 1];
    r = r < lo ? lo : (r > hi ? hi : r);
    out[(size_t)n * Vo + v] = (float)r;
}

// edge-pad by NPAD into the double-precision coefficient buffer
__global__ void __launch_bounds__(256) pad_kernel(const float *in, double *coef, Geo G)
{
    const int Dp = G.Di + 2 * NPAD, Hp = G.Hi + 2 * NPAD, Wp = G.Wi + 2 * NPAD;
    const long long Vp = (long long)Dp * Hp * Wp, Vi = (long long)G.Di * G.Hi * G.Wi;
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (v >= Vp) return;
    const int w = (int)(v % Wp), h = (int)((v / Wp) % Hp), d = (int)(v / ((long long)Wp * Hp));
    const int sd = min(max(d - NPAD, 0), G.Di - 1), sh = min(max(h - NPAD, 0), G.Hi - 1), sw = min(max(w - NPAD, 0), G.Wi - 1);
    coef[(size_t)n * Vp + v] = (double)in[(size_t)n * Vi + ((size_t)sd * G.Hi + sh) * G.Wi + sw];
}

// cubic B-spline prefilter along one axis, in place (ni_splines.c: _apply_filter_gain, _init_causal_reflect, the two
// recursions, _init_anticausal_reflect).  One thread per line; `len` samples `stride` apart.
__global__ void __launch_bounds__(128) prefilter_kernel(double *coef, long long nlines, int len, long long stride, long long inner,
                                                        long long outer_stride)
{
    const long long line = (long long)blockIdx.x * 128 + threadIdx.x;
    if (line >= nlines) return;
    // line -> (outer, inner): base = outer * outer_stride + inner
    double *c = coef + (line / inner) * outer_stride + (line % inner);
    const double z = sqrt(3.0) - 2.0;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    for (int i = 0; i < len; ++i) c[i * stride] *= gain;
    if (len < 2) return;
    const double z_n = pow(z, (double)len);
    const double c0 = c[0];
    double acc = c[0] + z_n * c[(len - 1) * stride];
    double z_i = z;
    for (int i = 1; i < len; ++i) {
        acc += z_i * (c[i * stride] + z_n * c[(len - 1 - i) * stride]);
        z_i *= z;
    }
    c[0] = acc * (z / (1.0 - z_n * z_n)) + c0;
    for (int i = 1; i < len; ++i) c[i * stride] += z * c[(i - 1) * stride];
    c[(len - 1) * stride] *= z / (z - 1.0);
    for (int i = len - 2; i >= 0; --i) c[i * stride] = z * (c[(i + 1) * stride] - c[i * stride]);
}

// order 3: start[axis][o] (first of four taps in the padded coefficient volume) and wgt[axis][o][4]
__global__ void __launch_bounds__(256) cubic_kernel(const double *coef, float *out, const int *start, const double *wgt, const float *mm, Geo G)
{
    const int Hp = G.Hi + 2 * NPAD, Wp = G.Wi + 2 * NPAD, Dp = G.Di + 2 * NPAD;
    const long long Vo = (long long)G.Do * G.Ho * G.Wo, Vp = (long long)Dp * Hp * Wp;
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (v >= Vo) return;
    const int w = (int)(v % G.Wo), h = (int)((v / G.Wo) % G.Ho), d = (int)(v / ((long long)G.Wo * G.Ho));
    const double *c = coef + (size_t)n * Vp;
    const int sd = start[d], sh = start[G.Do + h], sw = start[G.Do + G.Ho + w];
    const double *wd = wgt + 4 * d, *wh = wgt + 4 * (G.Do + h), *ww = wgt + 4 * (G.Do + G.Ho + w);
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int dd = min(max(sd + i, 0), Dp - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int hh = min(max(sh + j, 0), Hp - 1);
            const double wij = wd[i] * wh[j];
            const double *row = c + ((size_t)dd * Hp + hh) * Wp;
#pragma unroll
            for (int k = 0; k < 4; ++k) r += row[min(max(sw + k, 0), Wp - 1)] * (wij * ww[k]);
        }
    }
    const double lo = mm[2 * n], hi = mm[2 * n + 1];
    r = r < lo ? lo : (r > hi ? hi : r);
    out[(size_t)n * Vo + v] = (float)r;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Layout {
    size_t mm_off, idx_off, wgt_off, coef_off, total;
};

static Layout layout(int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int order)
{
    Layout L;
    const size_t nt = (size_t)Do + Ho + Wo;
    size_t off = 0;
    L.mm_off = off; off = align256(off + (size_t)N * 2 * sizeof(float));
    L.idx_off = off; off = align256(off + nt * sizeof(int));
    L.wgt_off = off; off = align256(off + nt * 4 * sizeof(double));
    L.coef_off = off;
    if (order == 3) off += align256((size_t)N * (Di + 2 * NPAD) * (Hi + 2 * NPAD) * (Wi + 2 * NPAD) * sizeof(double));
    L.total = off;
    return L;
}

// scipy NI_ZoomShift, grid_mode: cc = o; cc += 0.5; cc *= zoom; cc -= 0.5   (zoom = n_in / n_out, double)
static double coord(int o, int n_in, int n_out)
{
    volatile double zoom = n_out > 0 ? (double)n_in / (double)n_out : 1.0;
    volatile double cc = (double)o;
    cc = cc + 0.5;
    cc = cc * zoom;
    cc = cc - 0.5;
    return cc;
}

}  // namespace rsz

void preload_resize()
{
    DGTTA_TOUCH(rsz::minmax_kernel); DGTTA_TOUCH(rsz::gather_kernel<0>); DGTTA_TOUCH(rsz::gather_kernel<1>);
    DGTTA_TOUCH(rsz::pad_kernel); DGTTA_TOUCH(rsz::prefilter_kernel); DGTTA_TOUCH(rsz::cubic_kernel);
}

}  // namespace dgtta

using namespace dgtta;

extern "C" size_t dgtta_resize_edge_workspace_bytes(int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int order)
{
    if (N <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0) return 0;
    return rsz::layout(N, Di, Hi, Wi, Do, Ho, Wo, order).total;
}

extern "C" int dgtta_resize_edge(const float *in_dev, float *out_dev, int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                                 int order, void *workspace_dev, size_t workspace_bytes, dgtta_stream_t stream_)
{
    using namespace rsz;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!in_dev || !out_dev || !workspace_dev) { set_error("dgtta_resize_edge: null pointer"); return DGTTA_ENULL; }
    if (N <= 0 || N > 65535 || Di <= 0 || Hi <= 0 || Wi <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0) { set_error("dgtta_resize_edge: bad shape"); return DGTTA_EINVAL; }
    if (order != 0 && order != 1 && order != 3) { set_error("dgtta_resize_edge: order %d not in {0,1,3}", order); return DGTTA_EUNSUPPORTED; }
    const Layout L = layout(N, Di, Hi, Wi, Do, Ho, Wo, order);
    if (workspace_bytes < L.total || ((uintptr_t)workspace_dev & 255)) { set_error("dgtta_resize_edge: workspace too small or not 256-byte aligned"); return DGTTA_EWORKSPACE; }
    char *base = (char *)workspace_dev;
    float *mm = (float *)(base + L.mm_off);
    int *idx = (int *)(base + L.idx_off);
    double *wgt = (double *)(base + L.wgt_off);
    double *coef = (double *)(base + L.coef_off);
    const Geo G{N, Di, Hi, Wi, Do, Ho, Wo};

    // per-axis tables, evaluated in double with scipy's operation order
    const int nin[3] = {Di, Hi, Wi}, nout[3] = {Do, Ho, Wo};
    const size_t nt = (size_t)Do + Ho + Wo;
    std::vector<int> h_idx(nt);
    std::vector<double> h_w(nt * 4);    // order 3: four weights per output index; order 1: y = cc - floor(cc) in h_w[t]
    size_t t = 0;
    for (int a = 0; a < 3; ++a)
        for (int o = 0; o < nout[a]; ++o, ++t) {
            double cc = coord(o, nin[a], nout[a]);
            if (order == 3) {
                const double f = floor(cc), y = cc - f, z = 1.0 - y;
                h_idx[t] = (int)f - 1 + NPAD;
                double *w = &h_w[4 * t];
                w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
                w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
                w[0] = z * z * z / 6.0;
                w[3] = 1.0 - w[0] - w[1] - w[2];
            } else {
                cc = cc < 0.0 ? 0.0 : (cc > (double)(nin[a] - 1) ? (double)(nin[a] - 1) : cc);   // mode 'nearest'
                if (order == 0) h_idx[t] = (int)floor(cc + 0.5);
                else { const double f = floor(cc); h_idx[t] = (int)f; h_w[t] = cc - f; }
            }
        }
    cudaError_t e = cudaMemcpyAsync(idx, h_idx.data(), nt * sizeof(int), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && order == 3) e = cudaMemcpyAsync(wgt, h_w.data(), nt * 4 * sizeof(double), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && order == 1) e = cudaMemcpyAsync(wgt, h_w.data(), nt * sizeof(double), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { set_error("dgtta_resize_edge: table upload: %s", cudaGetErrorString(e)); return (int)e; }

    const long long Vi = (long long)Di * Hi * Wi, Vo = (long long)Do * Ho * Wo;
    minmax_kernel<<<N, 1024, 0, stream>>>(in_dev, Vi, mm);
    int rc = check_launch("resize minmax_kernel");
    if (rc) return rc;
    const dim3 ogrid((unsigned)((Vo + 255) / 256), (unsigned)N);
    if (order == 0) {
        gather_kernel<0><<<ogrid, 256, 0, stream>>>(in_dev, out_dev, idx, nullptr, mm, G);
        return check_launch("resize gather_kernel<0>");
    }
    if (order == 1) {
        gather_kernel<1><<<ogrid, 256, 0, stream>>>(in_dev, out_dev, idx, wgt, mm, G);
        return check_launch("resize gather_kernel<1>");
    }
    const int Dp = Di + 2 * NPAD, Hp = Hi + 2 * NPAD, Wp = Wi + 2 * NPAD;
    const long long Vp = (long long)Dp * Hp * Wp;
    pad_kernel<<<dim3((unsigned)((Vp + 255) / 256), (unsigned)N), 256, 0, stream>>>(in_dev, coef, G);
    rc = check_launch("resize pad_kernel");
    if (rc) return rc;
    // spline_filter runs axis 0, 1, 2 (scipy.ndimage.spline_filter)
    {
        const long long nl = (long long)N * Hp * Wp;      // axis D: lines indexed by (n, h*Wp + w)
        prefilter_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, stream>>>(coef, nl, Dp, (long long)Hp * Wp, (long long)Hp * Wp, Vp);
        rc = check_launch("resize prefilter_kernel D");
        if (rc) return rc;
    }
    {
        const long long nl = (long long)N * Dp * Wp;      // axis H: (n*Dp + d, w)
        prefilter_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, stream>>>(coef, nl, Hp, Wp, Wp, (long long)Hp * Wp);
        rc = check_launch("resize prefilter_kernel H");
        if (rc) return rc;
    }
    {
        const long long nl = (long long)N * Dp * Hp;      // axis W: (n*Dp*Hp + d*Hp + h, -)
        prefilter_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, stream>>>(coef, nl, Wp, 1, 1, Wp);
        rc = check_launch("resize prefilter_kernel W");
        if (rc) return rc;
    }
    cubic_kernel<<<ogrid, 256, 0, stream>>>(coef, out_dev, idx, wgt, mm, G);
    return check_launch("resize cubic_kernel");
}
