/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the DG-TTA input-transform hot path.
 * Nothing under oracle/ may be imported, linked or executed by the product (dg_tta_b200/);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * This header is a "template": dgtta_oracle.c includes it twice, once with REAL=float
 * (the reference's arithmetic type) and once with REAL=double (a higher-precision truth used to
 * judge both the reference and the CUDA path).  Every function cites the reference lines it
 * restates.  Layouts are the reference's: contiguous NCDHW float32.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against the fixtures in
 * the tests/golden npz files, which were produced by running the unmodified reference
 * (tests/golden/make_golden.py).
 */

#ifndef REAL
#error "include from dgtta_oracle.c"
#endif

#ifndef IMGT
#define IMGT float /* element type of oracle_mind_ssc's input image; double only for the *_f64x chain-truth variant */
#define IMGT_DEFAULTED
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static inline int FN(clampi)(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------------
 * MIND-SSC.  Reference: dg_tta/mind.py
 *   :104-136  the 12 (shift1, shift2) offset pairs (table below, verified against the dumped one-hot
 *             kernels in tests/golden/mind_shift_table.npz)
 *   :146-148  edge = I(clamp(p + delta*s1)) - I(clamp(p + delta*s2))   (ReplicationPad3d + dilated conv)
 *   :150-152  edge += randn_weighting * noise
 *   :153      ssd = smooth(edge**2, sigma): three replicate-padded 1-D passes, order D, H, W (:5-24,:39-41)
 *   :156      mind = ssd - min_c ssd
 *   :157      var = mean_c mind
 *   :158-160  var = clamp(var, 0.001*mean_all(var), 1000*mean_all(var))   (mean over B*D*H*W)
 *   :161-162  out = exp(-(mind / var))
 * ------------------------------------------------------------------------------------------------ */
static void FN(smooth_axis)(const REAL *src, REAL *dst, long planes, int D, int H, int W, int axis,
                            const REAL *taps, int ntaps)
{
    const int r = ntaps / 2;
    const long V = (long)D * H * W;
#pragma omp parallel for schedule(static)
    for (long pz = 0; pz < planes * D; ++pz) {
        const long p = pz / D;
        const int d = (int)(pz % D);
        const REAL *s = src + p * V;
        REAL *o = dst + p * V;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                REAL acc = 0;
                for (int t = 0; t < ntaps; ++t) {
                    int dd = d, hh = h, ww = w;
                    if (axis == 0) dd = FN(clampi)(d + t - r, 0, D - 1);
                    if (axis == 1) hh = FN(clampi)(h + t - r, 0, H - 1);
                    if (axis == 2) ww = FN(clampi)(w + t - r, 0, W - 1);
                    acc += taps[t] * s[((long)dd * H + hh) * W + ww];
                }
                o[((long)d * H + h) * W + w] = acc;
            }
    }
}

int FN(oracle_mind_ssc)(const IMGT *img, const float *noise, REAL *out, int B, int D, int H, int W,
                        int delta, const float *taps_f, int ntaps, float randn_weighting)
{
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || delta < 1 || ntaps < 1 || !(ntaps & 1)) return 1;
    const long V = (long)D * H * W;
    const long N = (long)B * 12 * V;
    REAL *a = (REAL *)malloc(sizeof(REAL) * N);
    REAL *b = (REAL *)malloc(sizeof(REAL) * N);
    REAL *taps = (REAL *)malloc(sizeof(REAL) * ntaps);
    if (!a || !b || !taps) { free(a); free(b); free(taps); return 2; }
    for (int t = 0; t < ntaps; ++t) taps[t] = (REAL)taps_f[t];
    const REAL rw = (REAL)randn_weighting;

    /* edge_selection ** 2 */
#pragma omp parallel for schedule(static)
    for (long bc = 0; bc < (long)B * 12; ++bc) {
        const int bi = (int)(bc / 12), c = (int)(bc % 12);
        const IMGT *I = img + bi * V;
        const int *s1 = MIND_SHIFT1[c], *s2 = MIND_SHIFT2[c];
        for (int d = 0; d < D; ++d)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const long q1 = ((long)FN(clampi)(d + delta * s1[0], 0, D - 1) * H +
                                     FN(clampi)(h + delta * s1[1], 0, H - 1)) * W +
                                    FN(clampi)(w + delta * s1[2], 0, W - 1);
                    const long q2 = ((long)FN(clampi)(d + delta * s2[0], 0, D - 1) * H +
                                     FN(clampi)(h + delta * s2[1], 0, H - 1)) * W +
                                    FN(clampi)(w + delta * s2[2], 0, W - 1);
                    const long p = ((long)d * H + h) * W + w;
                    REAL e = (REAL)I[q1] - (REAL)I[q2];
                    if (noise) e = e + rw * (REAL)noise[bc * V + p];
                    a[bc * V + p] = e * e;
                }
    }
    /* smooth: D, H, W */
    FN(smooth_axis)(a, b, (long)B * 12, D, H, W, 0, taps, ntaps);
    FN(smooth_axis)(b, a, (long)B * 12, D, H, W, 1, taps, ntaps);
    FN(smooth_axis)(a, b, (long)B * 12, D, H, W, 2, taps, ntaps);

    /* mind = ssd - min; var = mean_c(mind); global mean */
    REAL *var = a; /* reuse: B*V entries */
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (long bp = 0; bp < (long)B * V; ++bp) {
        const long bi = bp / V, p = bp % V;
        REAL *s = b + bi * 12 * V + p;
        REAL mn = s[0];
        for (int c = 1; c < 12; ++c) mn = s[c * V] < mn ? s[c * V] : mn;
        REAL sum = 0;
        for (int c = 0; c < 12; ++c) { s[c * V] -= mn; sum += s[c * V]; }
        var[bp] = sum / (REAL)12;
        total += (double)var[bp];
    }
    const REAL mean_all = (REAL)(total / (double)((long)B * V));
    const REAL lo = mean_all * (REAL)0.001, hi = mean_all * (REAL)1000;
#pragma omp parallel for schedule(static)
    for (long bp = 0; bp < (long)B * V; ++bp) {
        const long bi = bp / V, p = bp % V;
        REAL v = var[bp];
        /* torch.clamp(x, min, max) == min(max(x, min), max); NaN propagates */
        v = v < lo ? lo : v;
        v = v > hi ? hi : v;
        for (int c = 0; c < 12; ++c) {
            const REAL m = b[(bi * 12 + c) * V + p] / v;
            out[(bi * 12 + c) * V + p] = (REAL)EXPFN(-m);
        }
    }
    free(a); free(b); free(taps);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * GIN.  Reference: dg_tta/gin.py
 *   :94-107  per layer: ker [cout*B, cin, k,k,k], shift [cout*B]; grouped conv, groups=B, zero padding
 *            k//2, stride 1: sample b uses ker[b*cout:(b+1)*cout]
 *   :111-113 + shift; leaky_relu(0.01) unless last layer
 *   :139-164 channels: in -> interm -> ... -> interm -> in (N_LAYER layers)
 *   :197     mixed = alpha*net(x) + (1-alpha)*x
 *   :200-228 out = mixed * (1/(||mixed_b||_F + 1e-5)) * ||x_b||_F      (per sample)
 * params: for layer L in order: ker_L (contiguous [cout*B, cin, k,k,k]) followed by shift_L [cout*B].
 * ------------------------------------------------------------------------------------------------ */
int FN(oracle_gin)(const float *x, REAL *out, const float *params, const int *ksizes, const float *alphas,
                   int B, int D, int H, int W, int in_channels, int n_layer, int interm_channels)
{
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || n_layer < 2) return 1;
    const long V = (long)D * H * W;
    const int cmax = in_channels > interm_channels ? in_channels : interm_channels;
    REAL *cur = (REAL *)malloc(sizeof(REAL) * B * cmax * V);
    REAL *nxt = (REAL *)malloc(sizeof(REAL) * B * cmax * V);
    if (!cur || !nxt) { free(cur); free(nxt); return 2; }
    for (long i = 0; i < (long)B * in_channels * V; ++i) cur[i] = (REAL)x[i];
    int cin = in_channels;
    const float *pp = params;
    for (int L = 0; L < n_layer; ++L) {
        const int cout = (L == n_layer - 1) ? in_channels : interm_channels;
        const int k = ksizes[L], r = k / 2, k3 = k * k * k;
        const float *ker = pp;
        const float *shift = pp + (long)cout * B * cin * k3;
        pp = shift + (long)cout * B;
        const int act = (L != n_layer - 1);
#pragma omp parallel for schedule(static)
        for (long bod = 0; bod < (long)B * cout * D; ++bod) {
            const int d = (int)(bod % D);
            const int oc = (int)((bod / D) % cout);
            const int bi = (int)(bod / ((long)D * cout));
            const float *kk = ker + ((long)(bi * cout + oc) * cin) * k3;
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    REAL acc = 0;
                    for (int ic = 0; ic < cin; ++ic) {
                        const REAL *src = cur + ((long)bi * cin + ic) * V;
                        for (int a = 0; a < k; ++a) {
                            const int dd = d + a - r;
                            if (dd < 0 || dd >= D) continue;
                            for (int bq = 0; bq < k; ++bq) {
                                const int hh = h + bq - r;
                                if (hh < 0 || hh >= H) continue;
                                for (int c = 0; c < k; ++c) {
                                    const int ww = w + c - r;
                                    if (ww < 0 || ww >= W) continue;
                                    acc += (REAL)kk[ic * k3 + (a * k + bq) * k + c] *
                                           src[((long)dd * H + hh) * W + ww];
                                }
                            }
                        }
                    }
                    acc += (REAL)shift[bi * cout + oc];
                    if (act && acc < 0) acc *= (REAL)0.01;
                    nxt[((long)bi * cout + oc) * V + ((long)d * H + h) * W + w] = acc;
                }
        }
        REAL *t = cur; cur = nxt; nxt = t;
        cin = cout;
    }
    /* blend + Frobenius re-normalisation, per sample over all (in_channels) channels */
    const long CV = (long)in_channels * V;
    for (int bi = 0; bi < B; ++bi) {
        const REAL al = (REAL)alphas[bi];
        double s_in = 0.0, s_mix = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s_in, s_mix)
        for (long i = 0; i < CV; ++i) {
            const REAL xi = (REAL)x[bi * CV + i];
            const REAL m = al * cur[bi * CV + i] + ((REAL)1.0 - al) * xi;
            cur[bi * CV + i] = m;
            s_in += (double)xi * (double)xi;
            s_mix += (double)m * (double)m;
        }
        const REAL in_frob = (REAL)sqrt(s_in), self_frob = (REAL)sqrt(s_mix);
        const REAL inv = (REAL)1.0 / (self_frob + (REAL)1e-5);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < CV; ++i) out[bi * CV + i] = cur[bi * CV + i] * inv * in_frob;
    }
    free(cur); free(nxt);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Affine grid + grid_sample.  Reference call sites: dg_tta/tta/tta.py:143-147,523-532,548-551,
 * 571-575 and dg_tta/tta/torch_utils.py:55-73; the arithmetic is torch's (third party, pinned
 * torch==2.2.0 in poetry.lock; semantics followed: ATen/native/AffineGridGenerator.cpp
 * (linspace(-1,1,N)*(N-1)/N base grid, grid = base @ theta^T) and ATen/native/GridSampler.h
 * grid_sampler_unnormalize / clip_coordinates / within_bounds, align_corners=False).
 * theta: [B,3,4] rows produce (x,y,z) = (W,H,D)-axis coordinates.
 * mode: 0 trilinear, 1 nearest.  padding: 0 zeros, 1 border.
 * ------------------------------------------------------------------------------------------------ */
static inline REAL FN(base_coord)(int i, int n)
{
    /* torch.linspace(-1, 1, n)[i] * (n-1)/n.  linspace (ATen RangeFactories) is evaluated symmetrically from both ends
     * with one fused multiply-add per element; then a tensor multiply and a true division, each rounded.  Checked
     * bit for bit against torch (CPU build) for n = 2..399 while writing this oracle. */
    if (n <= 1) return (REAL)0;
    const REAL step = (REAL)2 / (REAL)(n - 1);
    const REAL v = (i < n / 2) ? FMAFN(step, (REAL)i, (REAL)-1) : FMAFN(-step, (REAL)(n - 1 - i), (REAL)1);
    return v * (REAL)(n - 1) / (REAL)n;
}

/* one row of grid = base_grid.view(N, DHW, 4).bmm(theta^T) (AffineGridGenerator.cpp): the K = 4 products accumulated
 * in order, the first rounded and the rest fused — the sequence the BLAS kernel executes (bit-exact vs torch's CPU
 * affine_grid on random affines and sizes, 0 differing elements) */
static inline REAL FN(grid_row)(const float *t, REAL xn, REAL yn, REAL zn)
{
    return FMAFN(zn, (REAL)t[2], FMAFN(yn, (REAL)t[1], xn * (REAL)t[0])) + (REAL)t[3];
}

static inline REAL FN(unnormalize)(REAL g, int size) { return ((g + (REAL)1) * (REAL)size - (REAL)1) / (REAL)2; }

static inline REAL FN(clip)(REAL v, int size)
{
    const REAL hi = (REAL)(size - 1);
    v = v > (REAL)0 ? v : (REAL)0; /* NaN -> 0 as in torch's min(max()) */
    return v < hi ? v : hi;
}

int FN(oracle_affine_sample)(const float *in, const float *theta, REAL *out, int B, int C, int Di, int Hi,
                             int Wi, int Do, int Ho, int Wo, int mode, int padding)
{
    const long Vi = (long)Di * Hi * Wi, Vo = (long)Do * Ho * Wo;
#pragma omp parallel for schedule(static)
    for (long bd = 0; bd < (long)B * Do; ++bd) {
        const int bi = (int)(bd / Do), d = (int)(bd % Do);
        const float *th = theta + bi * 12;
        for (int h = 0; h < Ho; ++h)
            for (int w = 0; w < Wo; ++w) {
                const REAL xn = FN(base_coord)(w, Wo), yn = FN(base_coord)(h, Ho), zn = FN(base_coord)(d, Do);
                REAL gx = FN(grid_row)(th + 0, xn, yn, zn);
                REAL gy = FN(grid_row)(th + 4, xn, yn, zn);
                REAL gz = FN(grid_row)(th + 8, xn, yn, zn);
                REAL ix = FN(unnormalize)(gx, Wi), iy = FN(unnormalize)(gy, Hi), iz = FN(unnormalize)(gz, Di);
                if (padding == 1) { ix = FN(clip)(ix, Wi); iy = FN(clip)(iy, Hi); iz = FN(clip)(iz, Di); }
                const long po = ((long)d * Ho + h) * Wo + w;
                if (mode == 1) {
                    const long x0 = (long)NEARBY(ix), y0 = (long)NEARBY(iy), z0 = (long)NEARBY(iz);
                    const int ok = x0 >= 0 && x0 < Wi && y0 >= 0 && y0 < Hi && z0 >= 0 && z0 < Di;
                    for (int c = 0; c < C; ++c)
                        out[((long)bi * C + c) * Vo + po] =
                            ok ? (REAL)in[((long)bi * C + c) * Vi + (z0 * Hi + y0) * Wi + x0] : (REAL)0;
                    continue;
                }
                const REAL fx = FLOORFN(ix), fy = FLOORFN(iy), fz = FLOORFN(iz);
                const long x0 = (long)fx, y0 = (long)fy, z0 = (long)fz;
                const REAL tx = ix - fx, ty = iy - fy, tz = iz - fz;
                for (int c = 0; c < C; ++c) {
                    const float *src = in + ((long)bi * C + c) * Vi;
                    REAL acc = 0;
                    for (int k = 0; k < 8; ++k) {
                        const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
                        const long xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
                        if (xx < 0 || xx >= Wi || yy < 0 || yy >= Hi || zz < 0 || zz >= Di) continue;
                        const REAL wgt = (dx ? tx : (REAL)1 - tx) * (dy ? ty : (REAL)1 - ty) * (dz ? tz : (REAL)1 - tz);
                        acc += (REAL)src[(zz * Hi + yy) * Wi + xx] * wgt;
                    }
                    out[((long)bi * C + c) * Vo + po] = acc;
                }
            }
    }
    return 0;
}

/* backward w.r.t. the input of the trilinear sampler (the adjoint scatter); needed by
 * tta.py:573-575 when have_grad_in covers the branch (tta.py:496-504).  grad_in must hold
 * B*C*Di*Hi*Wi entries; it is overwritten.  Serial over output voxels -> deterministic. */
int FN(oracle_affine_sample_bwd_input)(const float *grad_out, const float *theta, REAL *grad_in, int B, int C,
                                       int Di, int Hi, int Wi, int Do, int Ho, int Wo, int padding)
{
    const long Vi = (long)Di * Hi * Wi, Vo = (long)Do * Ho * Wo;
    for (long i = 0; i < (long)B * C * Vi; ++i) grad_in[i] = 0;
#pragma omp parallel for schedule(static)
    for (long bc = 0; bc < (long)B * C; ++bc) {
        const int bi = (int)(bc / C);
        const float *th = theta + bi * 12;
        REAL *gi = grad_in + bc * Vi;
        const float *go = grad_out + bc * Vo;
        for (int d = 0; d < Do; ++d)
            for (int h = 0; h < Ho; ++h)
                for (int w = 0; w < Wo; ++w) {
                    const REAL xn = FN(base_coord)(w, Wo), yn = FN(base_coord)(h, Ho), zn = FN(base_coord)(d, Do);
                    REAL gx = FN(grid_row)(th + 0, xn, yn, zn);
                    REAL gy = FN(grid_row)(th + 4, xn, yn, zn);
                    REAL gz = FN(grid_row)(th + 8, xn, yn, zn);
                    REAL ix = FN(unnormalize)(gx, Wi), iy = FN(unnormalize)(gy, Hi), iz = FN(unnormalize)(gz, Di);
                    if (padding == 1) { ix = FN(clip)(ix, Wi); iy = FN(clip)(iy, Hi); iz = FN(clip)(iz, Di); }
                    const REAL fx = FLOORFN(ix), fy = FLOORFN(iy), fz = FLOORFN(iz);
                    const long x0 = (long)fx, y0 = (long)fy, z0 = (long)fz;
                    const REAL tx = ix - fx, ty = iy - fy, tz = iz - fz;
                    const REAL g = (REAL)go[((long)d * Ho + h) * Wo + w];
                    for (int k = 0; k < 8; ++k) {
                        const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
                        const long xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
                        if (xx < 0 || xx >= Wi || yy < 0 || yy >= Hi || zz < 0 || zz >= Di) continue;
                        const REAL wgt = (dx ? tx : (REAL)1 - tx) * (dy ? ty : (REAL)1 - ty) * (dz ? tz : (REAL)1 - tz);
                        gi[(zz * Hi + yy) * Wi + xx] += g * wgt;
                    }
                }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Consistency loss of the TTA step.  Reference: dg_tta/tta/tta.py:263-269 and soft_dice_loss,
 * dg_tta/tta/torch_utils.py:90-104.
 *   :264-266  mask = (sum_c a > 0) * (sum_c b > 0)
 *   :267-268  sm_x = softmax_c(x) * mask
 *   torch_utils.py:94-95   nominator = mean_v 2 sm_a sm_b ; denominator = 0.5 mean_v (sm_a + sm_b)^2
 *   :97-102   dice = 1 if denominator.sum() == 0 else nominator / denominator     (per sample and class)
 *   tta.py:269  loss = 1 - mean(dice[:, start_class:])
 * Returns the loss; grad_a (may be NULL) receives d loss / d target_a (the mask carries no gradient).
 * ------------------------------------------------------------------------------------------------ */
REAL FN(oracle_consistency_loss)(const float *ta, const float *tb, REAL *grad_a, int B, int C, long V, int start_class)
{
    REAL *nom = (REAL *)calloc((size_t)B * C, sizeof(REAL)), *den = (REAL *)calloc((size_t)B * C, sizeof(REAL));
    REAL *pa = (REAL *)malloc((size_t)C * sizeof(REAL)), *pb = (REAL *)malloc((size_t)C * sizeof(REAL));
    for (int b = 0; b < B; ++b)
        for (long v = 0; v < V; ++v) {
            REAL sa = 0, sb = 0, ma = -INFINITY, mb = -INFINITY;
            for (int c = 0; c < C; ++c) {
                const REAL xa = ta[((long)b * C + c) * V + v], xb = tb[((long)b * C + c) * V + v];
                sa += xa; sb += xb;
                if (xa > ma) ma = xa;
                if (xb > mb) mb = xb;
            }
            if (!(sa > 0 && sb > 0)) continue;
            REAL ea = 0, eb = 0;
            for (int c = 0; c < C; ++c) {
                pa[c] = (REAL)EXPFN(ta[((long)b * C + c) * V + v] - ma); ea += pa[c];
                pb[c] = (REAL)EXPFN(tb[((long)b * C + c) * V + v] - mb); eb += pb[c];
            }
            for (int c = 0; c < C; ++c) {
                const REAL a = pa[c] / ea, bb = pb[c] / eb;
                nom[b * C + c] += 2 * a * bb;
                den[b * C + c] += (a + bb) * (a + bb);
            }
        }
    REAL dsum = 0;
    for (int i = 0; i < B * C; ++i) { nom[i] /= (REAL)V; den[i] = (REAL)0.5 * den[i] / (REAL)V; dsum += den[i]; }
    const int K = B * (C - start_class);
    REAL loss = 1;
    for (int b = 0; b < B; ++b)
        for (int c = start_class; c < C; ++c) loss -= (dsum == 0 ? (REAL)1 : nom[b * C + c] / den[b * C + c]) / (REAL)K;
    if (grad_a) {
        for (long i = 0; i < (long)B * C * V; ++i) grad_a[i] = 0;
        if (dsum != 0)
            for (int b = 0; b < B; ++b)
                for (long v = 0; v < V; ++v) {
                    REAL sa = 0, sb = 0, ma = -INFINITY, mb = -INFINITY;
                    for (int c = 0; c < C; ++c) {
                        const REAL xa = ta[((long)b * C + c) * V + v], xb = tb[((long)b * C + c) * V + v];
                        sa += xa; sb += xb;
                        if (xa > ma) ma = xa;
                        if (xb > mb) mb = xb;
                    }
                    if (!(sa > 0 && sb > 0)) continue;
                    REAL ea = 0, eb = 0;
                    for (int c = 0; c < C; ++c) {
                        pa[c] = (REAL)EXPFN(ta[((long)b * C + c) * V + v] - ma); ea += pa[c];
                        pb[c] = (REAL)EXPFN(tb[((long)b * C + c) * V + v] - mb); eb += pb[c];
                    }
                    REAL dot = 0;
                    for (int c = 0; c < C; ++c) {
                        pa[c] /= ea; pb[c] /= eb;
                        /* d loss / d sm_a[c]: loss = 1 - (1/K) sum_{c >= start} nom_c / den_c,
                         * nom_c = (1/V) sum 2 a b, den_c = (0.5/V) sum (a+b)^2 */
                        REAL q = 0;
                        if (c >= start_class) {
                            const REAL n = nom[b * C + c], dn = den[b * C + c];
                            q = -((REAL)1 / (REAL)K) * ((2 * pb[c] / (REAL)V) / dn - n * ((pa[c] + pb[c]) / (REAL)V) / (dn * dn));
                        }
                        pb[c] = q;
                        dot += pa[c] * q;
                    }
                    for (int c = 0; c < C; ++c) grad_a[((long)b * C + c) * V + v] = pa[c] * (pb[c] - dot);
                }
    }
    free(nom); free(den); free(pa); free(pb);
    return loss;
}

/* Label crop of get_batch.  Reference: dg_tta/tta/torch_utils.py:71-73 (grid_sample, mode="nearest", zeros padding, of
 * the one-hot label channels through the patch affine) and :79-82 (get_argmaxed_segs: background channel where the
 * labels sum to < 1, then argmax; lowest index wins ties). */
int FN(oracle_label_argmax)(const float *onehot, const float *theta, long long *out, int B, int L, int Di, int Hi, int Wi,
                            int Do, int Ho, int Wo)
{
    const long Vi = (long)Di * Hi * Wi, Vo = (long)Do * Ho * Wo;
    for (int b = 0; b < B; ++b) {
        const float *th = theta + b * 12;
        for (int d = 0; d < Do; ++d)
            for (int h = 0; h < Ho; ++h)
                for (int w = 0; w < Wo; ++w) {
                    const REAL xn = FN(base_coord)(w, Wo), yn = FN(base_coord)(h, Ho), zn = FN(base_coord)(d, Do);
                    const REAL gx = FN(grid_row)(th + 0, xn, yn, zn);
                    const REAL gy = FN(grid_row)(th + 4, xn, yn, zn);
                    const REAL gz = FN(grid_row)(th + 8, xn, yn, zn);
                    const REAL rx = NEARBY(FN(unnormalize)(gx, Wi)), ry = NEARBY(FN(unnormalize)(gy, Hi)),
                               rz = NEARBY(FN(unnormalize)(gz, Di));
                    long long label = 0;
                    if (rx >= 0 && rx < Wi && ry >= 0 && ry < Hi && rz >= 0 && rz < Di) {
                        const long off = ((long)rz * Hi + (long)ry) * Wi + (long)rx;
                        REAL sum = 0, best = -INFINITY;
                        int arg = 0;
                        for (int l = 0; l < L; ++l) {
                            const REAL v = onehot[((long)b * L + l) * Vi + off];
                            sum += v;
                            if (v > best) { best = v; arg = l + 1; }
                        }
                        const REAL bg = sum < 1 ? (REAL)1 : (REAL)0;
                        label = bg >= best ? 0 : arg;
                    }
                    out[(long)b * Vo + ((long)d * Ho + h) * Wo + w] = label;
                }
    }
    return 0;
}

#ifdef IMGT_DEFAULTED
#undef IMGT
#undef IMGT_DEFAULTED
#endif
#undef FN
#undef CAT
#undef CAT_
