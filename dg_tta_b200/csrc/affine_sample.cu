// Affine view warp: F.affine_grid + F.grid_sample fused into one gather (forward) / scatter (adjoint).
//
// Replaces the op pairs at dg_tta/tta/tta.py:523-532+548-551 (image warp, border padding),
// tta.py:571-575 (inverse warp of the prediction, zeros padding, autograd w.r.t. the input) and
// dg_tta/tta/torch_utils.py:55-73 (patch crop; nearest for labels).  align_corners=False throughout.
// Coordinates follow torch (third party; ATen/native/AffineGridGenerator.cpp, GridSampler.h):
//     base_i = linspace(-1, 1, N)[i] * (N-1) / N          linspace: fma(step, i, -1) below N/2, fma(-step, N-1-i, 1) above
//     (gx,gy,gz) = base @ theta[b]^T                      K = 4 accumulated in order with FMAs: fma(z,t2, fma(y,t1, x*t0)) + t3
//   — the rounding sequence torch's kernels execute (RangeFactories linspace, tensor mul / true div, bmm), reproduced
//   operation by operation so that mode="nearest" picks the same source voxel on .5 ties: bit-exact index work
//   (tests/test_sampler_gpu.py compares the label crops with np.array_equal against the reference's outputs).
//     ix = ((gx + 1) * W_in - 1) / 2 ; border: clamp to [0, W_in-1] ; zeros: corners outside contribute 0
// No grid tensor is ever materialised (the reference builds three 12 B/voxel grids per warp).
#include "sampler.cuh"
#include <atomic>

namespace dgtta {

// grid: x = w-tiles * h-tiles, y = d, z = b.  Channels are looped inside (4 at a time: 32 independent gathers in
// flight) so that coordinates and weights are computed once per voxel.
template <int INTERP, int PAD>
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_sample_fwd_kernel(const __grid_constant__ SampleParams P)
{
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, b, w, h, d);
    const float *src = P.in + (size_t)b * P.C * Vi;
    float *dst = P.out + (size_t)b * P.C * Vo + ((size_t)d * P.Ho + h) * P.Wo + w;
    if (INTERP == DGTTA_INTERP_NEAREST) {
        const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
        const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
        const size_t off = ok ? ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx : 0;
        for (int ch = 0; ch < P.C; ++ch) __stcs(dst + ch * Vo, ok ? __ldg(src + ch * Vi + off) : 0.f);
        return;
    }
    Corners q = trilinear_corners(P, c);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (q.off[k] < 0) { q.off[k] = 0; q.wgt[k] = 0.f; }
    if (P.bias) {
        // get_batch's image crop (torch_utils.py:58-62): grid_sample(vol - min, zeros) + min in one gather — the shift is
        // applied to every in-bounds corner before the weighting, the out-of-bounds corners contribute 0 (their weight is 0)
        const float bias = __ldg(P.bias + b);
        for (int ch = 0; ch < P.C; ++ch) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(__fsub_rn(__ldg(src + (size_t)ch * Vi + q.off[k]), bias), q.wgt[k], acc);
            __stcs(dst + (size_t)ch * Vo, __fadd_rn(acc, bias));
        }
        return;
    }
    // channel loop: the eight corner offsets and weights stay in registers, only the two channel base pointers move
    // (one 64-bit add each per trip); four channels per trip keep 32 independent gathers in flight
    const float *cb = src;
    float *db = dst;
    int ch = 0;
    for (; ch + 4 <= P.C; ch += 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = __ldg(cb + (size_t)u * Vi + q.off[k]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(v[u][k], q.wgt[k], acc);
            __stcs(db + (size_t)u * Vo, acc);
        }
        cb += 4 * Vi;
        db += 4 * Vo;
    }
    for (; ch + 1 < P.C; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(cb + q.off[k]), q.wgt[k], acc);
        __stcs(db, acc);
        cb += Vi;
        db += Vo;
    }
    if (ch < P.C) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(cb + q.off[k]), q.wgt[k], acc);
        __stcs(db, acc);
    }
}

// adjoint of the trilinear forward: scatter grad_out into grad_in (pre-zeroed) with red.global.add.
// A warp is 32 consecutive output voxels along w; under the near-identity affines of the TTA loop lane i's x0+1 corner
// is lane i+1's x0 corner, so that half of each lane's contributions is handed to the neighbour by shuffle and added
// there first: ~4.1 instead of 8 global reductions per voxel and channel (the scatter is bound by L2 atomics).
template <int PAD>
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_sample_bwd_kernel(const __grid_constant__ SampleParams P, const int *skip_if_set)
{
    if (skip_if_set && *skip_if_set) return;   // the deterministic gather handles this call
    // here P.in = grad_out [B,C,Do,Ho,Wo], P.out = grad_in [B,C,Di,Hi,Wi]
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x;   // blockDim.x == 32: a warp is one row of the block
    const int w = w0 + lane, h = h0 + threadIdx.y;
    const bool active = w < P.Wo && h < P.Ho;   // inactive lanes stay for the shuffles and contribute nothing
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, b, w, h, d);
    Corners q = trilinear_corners(P, c);
    if (!active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) q.off[k] = -1;
    }
    // pair p = (dy, dz): corners 2p (x0) and 2p+1 (x0+1)
    bool give[4], take[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int next_x0 = __shfl_down_sync(0xffffffffu, q.off[2 * p], 1);
        give[p] = lane < 31 && q.off[2 * p + 1] >= 0 && next_x0 == q.off[2 * p + 1];
        take[p] = __shfl_up_sync(0xffffffffu, (int)give[p], 1) != 0 && lane > 0;
    }
    const float *go = P.in + (size_t)b * P.C * Vo + ((size_t)d * P.Ho + min(h, P.Ho - 1)) * P.Wo + min(w, P.Wo - 1);
    float *gi = P.out + (size_t)b * P.C * Vi;
    auto scatter = [&](float g, float *base) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float v1 = g * q.wgt[2 * p + 1];
            const float from_left = __shfl_up_sync(0xffffffffu, v1, 1);
            const float v0 = g * q.wgt[2 * p] + (take[p] ? from_left : 0.f);
            if (q.off[2 * p] >= 0) atomicAdd(base + q.off[2 * p], v0);
            if (q.off[2 * p + 1] >= 0 && !give[p]) atomicAdd(base + q.off[2 * p + 1], v1);
        }
    };
    int ch = 0;
    for (; ch + 4 <= P.C; ch += 4) {
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = active ? __ldg(go + (size_t)u * Vo) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) scatter(g[u], gi + (size_t)u * Vi);
        go += 4 * Vo;
        gi += 4 * Vi;
    }
    for (; ch < P.C; ++ch) {
        scatter(active ? __ldg(go) : 0.f, gi);
        go += Vo;
        gi += Vi;
    }
}

// Deterministic adjoint (zeros padding): a GATHER over the source voxels instead of the scatter above.  A thread owns one
// source voxel s of one sample, enumerates the output voxels whose trilinear footprint contains s and accumulates
// w(o, s) * grad_out[c][o] for (up to) 16 channels in registers, in a fixed order — no atomics, no zero-init pass, every
// grad_in element written exactly once, bit-identical from run to run.
//   * Candidates.  In index space the forward is (up to float rounding) affine, p = M o + c with
//     M[i][j] = theta[i][j] S_i / N_j,  c_i = S_i/2 (sum_j theta[i][j] (1/N_j - 1) + theta[i][3] + 1) - 1/2
//     (S = input size, N = output size, both in x, y, z order).  s receives from o iff p(o) lies in [s-1, s+1) on every
//     axis, so o lies in M^-1 of a cube of half-side 1 around s: the box  M^-1 (s - c) +- |M^-1| (1.01, 1.01, 1.01)
//     (row-wise L1 norms; the 1 % covers the rounding gap between this approximation and the exact coordinates).
//   * Decision.  Every candidate is then tested with the EXACT coordinates of the forward (source_coords, same
//     operations in the same order), corner by corner: x0 = floor(ix) must equal sx or sx - 1, etc., and the weight is the
//     forward's expression.  The approximation only limits the search; it never decides.
//   * Near-identity affines (the TTA loop's) give 3 x 3 x 3 ... 4 x 4 x 4 candidates of which ~8 contribute.  An affine
//     that magnifies strongly makes the box large: bwd_plan_kernel measures the box per sample first and the launch
//     falls back to the scatter for the whole call when any box edge exceeds GATHER_MAX_EDGE (the plan kernel writes a flag
//     both paths read: the choice is made on the device, no host synchronisation).
constexpr int GATHER_CH = 16;
constexpr int GATHER_MAX_EDGE = 6;
constexpr int PLAN_SLOTS = 64;
__device__ int g_bwd_plan[PLAN_SLOTS];   // 1: use the gather, 0: use the scatter (one slot per call in flight, round robin)

struct IndexAffine {
    float Mi[3][3];   // M^-1
    float c[3];
    float r[3];       // half-extent of the candidate box per output axis (w, h, d)
    bool ok;
};

__device__ __forceinline__ IndexAffine index_affine(const SampleParams &P, int b)
{
    const float *th = P.theta + b * 12;
    const float S[3] = {(float)P.Wi, (float)P.Hi, (float)P.Di}, N[3] = {(float)P.Wo, (float)P.Ho, (float)P.Do};
    float M[3][3];
    IndexAffine A;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float acc = __ldg(th + 4 * i + 3) + 1.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float t = __ldg(th + 4 * i + j);
            M[i][j] = t * S[i] / N[j];
            acc += t * (1.f / N[j] - 1.f);
        }
        A.c[i] = 0.5f * S[i] * acc - 0.5f;
    }
    const float c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1], c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2],
                c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const float det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
    A.ok = fabsf(det) > 1e-12f && det == det;
    const float id = A.ok ? 1.f / det : 0.f;
    A.Mi[0][0] = c00 * id; A.Mi[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id; A.Mi[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    A.Mi[1][0] = c01 * id; A.Mi[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; A.Mi[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    A.Mi[2][0] = c02 * id; A.Mi[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id; A.Mi[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
#pragma unroll
    for (int j = 0; j < 3; ++j) A.r[j] = 1.01f * (fabsf(A.Mi[j][0]) + fabsf(A.Mi[j][1]) + fabsf(A.Mi[j][2]));
    return A;
}

__global__ void bwd_plan_kernel(const __grid_constant__ SampleParams P, int slot)
{
    // one thread: the gather is used iff every sample's candidate box has at most GATHER_MAX_EDGE voxels per edge
    bool gather = true;
    for (int b = 0; b < P.B; ++b) {
        const IndexAffine A = index_affine(P, b);
        gather = gather && A.ok && 2.f * A.r[0] + 1.f <= (float)GATHER_MAX_EDGE && 2.f * A.r[1] + 1.f <= (float)GATHER_MAX_EDGE &&
                 2.f * A.r[2] + 1.f <= (float)GATHER_MAX_EDGE;
    }
    g_bwd_plan[slot] = gather ? 1 : 0;
}

__global__ void __launch_bounds__(256) bwd_zero_kernel(float4 *p, size_t n4, float *tail, int ntail, int slot)
{
    if (g_bwd_plan[slot]) return;   // the gather writes every element itself
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

#ifndef DGTTA_GATHER_MINB
#define DGTTA_GATHER_MINB 2
#endif
__global__ void __launch_bounds__(SAMPLE_THREADS, DGTTA_GATHER_MINB) affine_sample_bwd_gather_kernel(const __grid_constant__ SampleParams P, int slot)
{
    // P.in = grad_out [B,C,Do,Ho,Wo], P.out = grad_in [B,C,Di,Hi,Wi]; grid over SOURCE voxels: x = w-tiles * h-tiles, y = z, z = b
    if (!g_bwd_plan[slot]) return;
    const int ntw = (P.Wi + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int sx = tw * SBX + threadIdx.x, sy = th_ * SBY + threadIdx.y, sz = blockIdx.y, b = blockIdx.z;
    if (sx >= P.Wi || sy >= P.Hi) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const IndexAffine A = index_affine(P, b);
    const float dx_ = (float)sx - A.c[0], dy_ = (float)sy - A.c[1], dz_ = (float)sz - A.c[2];
    int lo[3], hi[3];
    const int N[3] = {P.Wo, P.Ho, P.Do};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float oc = A.Mi[j][0] * dx_ + A.Mi[j][1] * dy_ + A.Mi[j][2] * dz_;
        lo[j] = max(0, (int)ceilf(oc - A.r[j]));
        hi[j] = min(N[j] - 1, (int)floorf(oc + A.r[j]));
    }
    float t[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) t[k] = __ldg(P.theta + b * 12 + k);
    const float *go = P.in + (size_t)b * P.C * Vo;
    float *gi = P.out + (size_t)b * P.C * Vi + ((size_t)sz * P.Hi + sy) * P.Wi + sx;
    for (int c0 = 0; c0 < P.C; c0 += GATHER_CH) {
        const int nc = min(GATHER_CH, P.C - c0);
        float acc[GATHER_CH];
#pragma unroll
        for (int c = 0; c < GATHER_CH; ++c) acc[c] = 0.f;
        for (int od = lo[2]; od <= hi[2]; ++od) {
            const float zn = base_coord(od, P.Do, P.step_d, P.rcp_d, P.exact_div);
            for (int oh = lo[1]; oh <= hi[1]; ++oh) {
                const float yn = base_coord(oh, P.Ho, P.step_h, P.rcp_h, P.exact_div);
                for (int ow = lo[0]; ow <= hi[0]; ++ow) {
                    const float xn = base_coord(ow, P.Wo, P.step_w, P.rcp_w, P.exact_div);
                    // the forward's coordinates, operation by operation (source_coords / trilinear_corners)
                    const float gz = __fadd_rn(__fmaf_rn(zn, t[10], __fmaf_rn(yn, t[9], __fmul_rn(xn, t[8]))), t[11]);
                    const float iz = unnormalize(gz, P.Di);
                    const float fz = floorf(iz);
                    const int ez = sz - (int)fminf(fmaxf(fz, -2.f), (float)P.Di);
                    if ((unsigned)ez > 1u) continue;
                    const float gy = __fadd_rn(__fmaf_rn(zn, t[6], __fmaf_rn(yn, t[5], __fmul_rn(xn, t[4]))), t[7]);
                    const float iy = unnormalize(gy, P.Hi);
                    const float fy = floorf(iy);
                    const int ey = sy - (int)fminf(fmaxf(fy, -2.f), (float)P.Hi);
                    if ((unsigned)ey > 1u) continue;
                    const float gx = __fadd_rn(__fmaf_rn(zn, t[2], __fmaf_rn(yn, t[1], __fmul_rn(xn, t[0]))), t[3]);
                    const float ix = unnormalize(gx, P.Wi);
                    const float fx = floorf(ix);
                    const int ex = sx - (int)fminf(fmaxf(fx, -2.f), (float)P.Wi);
                    if ((unsigned)ex > 1u) continue;
                    const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
                    const float wgt = (ex ? tx : 1.f - tx) * (ey ? ty : 1.f - ty) * (ez ? tz : 1.f - tz);
                    const float *g = go + (size_t)c0 * Vo + ((size_t)od * P.Ho + oh) * P.Wo + ow;
#pragma unroll
                    for (int c = 0; c < GATHER_CH; ++c)
                        if (c < nc) acc[c] = fmaf(wgt, __ldg(g + (size_t)c * Vo), acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < GATHER_CH; ++c)
            if (c < nc) __stcs(gi + (size_t)(c0 + c) * Vi, acc[c]);
    }
}

// Label crop of get_batch (dg_tta/tta/torch_utils.py:71-82): nearest-neighbour sampling of the one-hot label channels
// (zeros outside), background channel prepended where no label is set, argmax over channels — in one pass, writing the
// int64 label map directly instead of an L-channel float patch plus sum / cat / argmax passes.  Ties go to the lowest
// index like torch.argmax; the background channel is index 0.
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_label_argmax_kernel(const __grid_constant__ SampleParams P, long long *out)
{
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, b, w, h, d);
    const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
    const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
    long long label = 0;
    if (ok) {
        const float *src = P.in + (size_t)b * P.C * Vi + ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx;
        float sum = 0.f, best = -__int_as_float(0x7f800000);
        int arg = 0;
        for (int ch = 0; ch < P.C; ++ch) {
            const float v = __ldg(src + (size_t)ch * Vi);
            sum += v;
            if (v > best) { best = v; arg = ch + 1; }
        }
        const float bg = sum < 1.0f ? 1.f : 0.f;     // get_argmaxed_segs: (segs.sum(1) < 1.0).float()
        label = bg >= best ? 0 : arg;
    }
    // outside the volume every channel samples 0: sum = 0 < 1 -> background
    out[(size_t)b * Vo + ((size_t)d * P.Ho + h) * P.Wo + w] = label;
}

// Integer label path.  Nearest sampling picks ONE source voxel, and get_argmaxed_segs is a per-voxel
// function of that voxel's L channel values, so the two commute: argmaxed(nearest_sample(onehot)) ==
// nearest_sample(argmaxed(onehot)) with 0 (background) outside the volume.  label_map_kernel evaluates the per-voxel
// rule once per volume (the host side caches it on the resident tensor); label_gather_kernel then crops from a
// 2 B/voxel map instead of the L x 4 B/voxel one-hot volume (104 channels in the reference's TotalSegmentator tasks,
// dg_tta/tta/nnunet_utils.py:191-195).
__global__ void __launch_bounds__(256) label_map_kernel(const float *onehot, short *map, int L, long long V)
{
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (v >= V) return;
    const float *src = onehot + (size_t)b * L * V + v;
    float sum = 0.f, best = -__int_as_float(0x7f800000);
    int arg = 0;
    for (int ch = 0; ch < L; ++ch) {
        const float x = __ldg(src + (size_t)ch * V);
        sum += x;                                   // torch's sum over dim 1 runs in channel order too
        if (x > best) { best = x; arg = ch + 1; }
    }
    const float bg = sum < 1.0f ? 1.f : 0.f;
    map[(size_t)b * V + v] = (short)(bg >= best ? 0 : arg);
}

__global__ void __launch_bounds__(SAMPLE_THREADS) label_gather_kernel(const __grid_constant__ SampleParams P, const short *map,
                                                                      long long *out)
{
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w = tw * SBX + threadIdx.x, h = th_ * SBY + threadIdx.y, d = blockIdx.y, b = blockIdx.z;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, b, w, h, d);
    const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
    const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
    long long label = 0;
    if (ok) label = map[(size_t)b * Vi + ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx];
    out[(size_t)b * Vo + ((size_t)d * P.Ho + h) * P.Wo + w] = label;
}

// min over a volume (img.min() of torch_utils.py:58): warp-shuffle + block reduction to per-block partials, then one
// block over the partials.  min is exact in any order, so the result is bitwise torch's.
constexpr int MIN_BLOCKS = 592;   // 148 SMs x 4
__global__ void __launch_bounds__(256) volume_min_partial_kernel(const float *in, long long n, float *partial)
{
    __shared__ float red[8];
    float m = __int_as_float(0x7f800000);
    bool nan = false;
    const long long n4 = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) ? n / 4 : 0;
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = __ldg(in4 + i);
        m = fminf(fminf(m, v.x), fminf(fminf(v.y, v.z), v.w));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = __ldg(in + i);
        m = fminf(m, v);
        nan |= v != v;
    }
    if (nan) m = __int_as_float(0x7fc00000);      // torch.min propagates NaN; fminf would drop it
    // NaN-propagating reduction: a NaN partial wins
    auto comb = [](float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? red[threadIdx.x] : __int_as_float(0x7f800000);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) partial[blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(1024) volume_min_final_kernel(const float *partial, int n, float *out)
{
    __shared__ float red[32];
    auto comb = [](float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); };
    float m = threadIdx.x < n ? partial[threadIdx.x] : __int_as_float(0x7f800000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[0] = m;
    }
}

void preload_sampler()
{
    DGTTA_TOUCH(label_map_kernel);
    DGTTA_TOUCH(label_gather_kernel);
    DGTTA_TOUCH(volume_min_partial_kernel);
    DGTTA_TOUCH(volume_min_final_kernel);
    DGTTA_TOUCH(affine_label_argmax_kernel);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_BORDER>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_BORDER>);
    DGTTA_TOUCH(affine_sample_bwd_kernel<DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_bwd_kernel<DGTTA_PAD_BORDER>);
}

static int sample_check(const void *a, const void *t, const void *o, int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo)
{
    if (!a || !t || !o) { set_error("dgtta_affine_sample: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || C <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0 || B > 65535 || Do > 65535) {
        set_error("dgtta_affine_sample: bad shape");
        return DGTTA_EINVAL;
    }
    if ((size_t)Di * Hi * Wi >= (size_t)1 << 31 || (size_t)Do * Ho * Wo >= (size_t)1 << 31) {
        set_error("dgtta_affine_sample: volume too large");
        return DGTTA_EINVAL;
    }
    return 0;
}

static dim3 sample_grid(int B, int Do, int Ho, int Wo)
{
    return dim3((unsigned)(((Wo + SBX - 1) / SBX) * ((Ho + SBY - 1) / SBY)), (unsigned)Do, (unsigned)B);
}

}  // namespace dgtta

using namespace dgtta;

static int sample_fwd(const float *in_dev, const float *theta_dev, const float *bias_dev, float *out_dev, int B, int C, int Di,
                      int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding, dgtta_stream_t stream_);

extern "C" int dgtta_affine_sample_fwd(const float *in_dev, const float *theta_dev, float *out_dev, int B, int C,
                                       int Di, int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding,
                                       dgtta_stream_t stream_)
{
    return sample_fwd(in_dev, theta_dev, nullptr, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo, interp, padding, stream_);
}

extern "C" int dgtta_affine_crop_shifted_fwd(const float *in_dev, const float *theta_dev, const float *shift_dev, float *out_dev,
                                             int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                                             dgtta_stream_t stream_)
{
    if (!shift_dev) { set_error("dgtta_affine_crop_shifted_fwd: null shift"); return DGTTA_ENULL; }
    return sample_fwd(in_dev, theta_dev, shift_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo, DGTTA_INTERP_TRILINEAR,
                      DGTTA_PAD_ZEROS, stream_);
}

static int sample_fwd(const float *in_dev, const float *theta_dev, const float *bias_dev, float *out_dev, int B, int C, int Di,
                      int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding, dgtta_stream_t stream_)
{
    int rc = sample_check(in_dev, theta_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    if ((interp != DGTTA_INTERP_TRILINEAR && interp != DGTTA_INTERP_NEAREST) ||
        (padding != DGTTA_PAD_ZEROS && padding != DGTTA_PAD_BORDER)) {
        set_error("dgtta_affine_sample_fwd: bad interp/padding");
        return DGTTA_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const SampleParams P = make_sample_params(in_dev, theta_dev, out_dev, bias_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    const dim3 grid = sample_grid(B, Do, Ho, Wo);
    const dim3 block(SBX, SBY, 1);
    if (interp == DGTTA_INTERP_TRILINEAR) {
        if (padding == DGTTA_PAD_ZEROS) affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P);
        else affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P);
    } else {
        if (padding == DGTTA_PAD_ZEROS) affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P);
        else affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P);
    }
    return check_launch("affine_sample_fwd_kernel");
}

extern "C" int dgtta_affine_sample_bwd_input(const float *grad_out_dev, const float *theta_dev, float *grad_in_dev,
                                             int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                                             int padding, dgtta_stream_t stream_)
{
    int rc = sample_check(grad_out_dev, theta_dev, grad_in_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    if (padding != DGTTA_PAD_ZEROS && padding != DGTTA_PAD_BORDER) { set_error("dgtta_affine_sample_bwd_input: bad padding"); return DGTTA_EINVAL; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const SampleParams P = make_sample_params(grad_out_dev, theta_dev, grad_in_dev, nullptr, B, C, Di, Hi, Wi, Do, Ho, Wo);
    const dim3 grid = sample_grid(B, Do, Ho, Wo);
    const dim3 block(SBX, SBY, 1);
    const size_t n = (size_t)B * C * Di * Hi * Wi;
    if (padding == DGTTA_PAD_ZEROS && getenv("DGTTA_SAMPLE_BWD_DETERMINISTIC") != nullptr) {
        // opt-in (DGTTA_SAMPLE_BWD_DETERMINISTIC=1): deterministic gather — bit-identical gradients from run to run at ~2.4x
        // the scatter's time (1.12 vs 0.47 ms on 2 x 14 x 128^3) — unless the plan kernel finds an affine that magnifies too
        // much for it (then: zero + scatter); the decision lives in a device flag, so four launches are queued and two of
        // them return at once
        static std::atomic<unsigned> next_slot{0};
        const int slot = (int)(next_slot.fetch_add(1) % PLAN_SLOTS);
        int *flag = nullptr;
        if (cudaGetSymbolAddress((void **)&flag, g_bwd_plan) != cudaSuccess) { set_error("dgtta_affine_sample_bwd_input: plan symbol"); return (int)cudaGetLastError(); }
        bwd_plan_kernel<<<1, 1, 0, stream>>>(P, slot);
        int rc2 = check_launch("bwd_plan_kernel");
        if (rc2) return rc2;
        affine_sample_bwd_gather_kernel<<<sample_grid(B, Di, Hi, Wi), block, 0, stream>>>(P, slot);
        rc2 = check_launch("affine_sample_bwd_gather_kernel");
        if (rc2) return rc2;
        const bool al = (reinterpret_cast<uintptr_t>(grad_in_dev) & 15) == 0;
        const size_t n4 = al ? n / 4 : 0;
        size_t zb = (n4 + 255) / 256;
        if (zb > (size_t)sm_count() * 8) zb = (size_t)sm_count() * 8;
        if (zb < 1) zb = 1;
        if (!al || n - 4 * n4 > 256) {   // odd alignment: plain memset-style pass is not worth a second kernel shape
            // (never the case for torch allocations; keep the semantics with the scatter path alone)
            cudaError_t e = cudaMemsetAsync(grad_in_dev, 0, n * sizeof(float), stream);
            if (e != cudaSuccess) { set_error("dgtta_affine_sample_bwd_input: memset: %s", cudaGetErrorString(e)); return (int)e; }
            affine_sample_bwd_kernel<DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P, nullptr);
            return check_launch("affine_sample_bwd_kernel");
        }
        bwd_zero_kernel<<<(unsigned)zb, 256, 0, stream>>>(reinterpret_cast<float4 *>(grad_in_dev), n4, grad_in_dev + 4 * n4, (int)(n - 4 * n4), slot);
        rc2 = check_launch("bwd_zero_kernel");
        if (rc2) return rc2;
        affine_sample_bwd_kernel<DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P, flag + slot);
        return check_launch("affine_sample_bwd_kernel");
    }
    cudaError_t e = cudaMemsetAsync(grad_in_dev, 0, n * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("dgtta_affine_sample_bwd_input: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (padding == DGTTA_PAD_ZEROS) affine_sample_bwd_kernel<DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P, nullptr);
    else affine_sample_bwd_kernel<DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P, nullptr);
    return check_launch("affine_sample_bwd_kernel");
}

extern "C" int dgtta_affine_label_argmax(const float *onehot_dev, const float *theta_dev, long long *out_dev, int B, int L,
                                         int Di, int Hi, int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream_)
{
    int rc = sample_check(onehot_dev, theta_dev, out_dev, B, L, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    const SampleParams P = make_sample_params(onehot_dev, theta_dev, nullptr, nullptr, B, L, Di, Hi, Wi, Do, Ho, Wo);
    affine_label_argmax_kernel<<<sample_grid(B, Do, Ho, Wo), dim3(SBX, SBY, 1), 0, (cudaStream_t)stream_>>>(P, out_dev);
    return check_launch("affine_label_argmax_kernel");
}

extern "C" int dgtta_label_map_from_onehot(const float *onehot_dev, short *map_dev, int B, int L, long long V,
                                           dgtta_stream_t stream_)
{
    if (!onehot_dev || !map_dev) { set_error("dgtta_label_map_from_onehot: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || B > 65535 || L <= 0 || L > 32766 || V <= 0 || V >= (1ll << 38)) { set_error("dgtta_label_map_from_onehot: bad shape"); return DGTTA_EINVAL; }
    label_map_kernel<<<dim3((unsigned)((V + 255) / 256), (unsigned)B), 256, 0, (cudaStream_t)stream_>>>(onehot_dev, map_dev, L, V);
    return check_launch("label_map_kernel");
}

extern "C" int dgtta_affine_label_gather(const short *map_dev, const float *theta_dev, long long *out_dev, int B, int Di, int Hi,
                                         int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream_)
{
    int rc = sample_check(map_dev, theta_dev, out_dev, B, 1, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    const SampleParams P = make_sample_params(nullptr, theta_dev, nullptr, nullptr, B, 1, Di, Hi, Wi, Do, Ho, Wo);
    label_gather_kernel<<<sample_grid(B, Do, Ho, Wo), dim3(SBX, SBY, 1), 0, (cudaStream_t)stream_>>>(P, map_dev, out_dev);
    return check_launch("label_gather_kernel");
}

extern "C" size_t dgtta_volume_min_workspace_bytes(void) { return MIN_BLOCKS * sizeof(float); }

extern "C" int dgtta_volume_min(const float *in_dev, long long numel, float *out_dev, void *workspace_dev, size_t workspace_bytes,
                                dgtta_stream_t stream_)
{
    if (!in_dev || !out_dev || !workspace_dev) { set_error("dgtta_volume_min: null pointer"); return DGTTA_ENULL; }
    if (numel <= 0) { set_error("dgtta_volume_min: empty input"); return DGTTA_EINVAL; }
    if (workspace_bytes < MIN_BLOCKS * sizeof(float)) { set_error("dgtta_volume_min: workspace too small"); return DGTTA_EWORKSPACE; }
    const long long want = (numel + 256 * 16 - 1) / (256 * 16);
    const int blocks = (int)(want < 1 ? 1 : (want > MIN_BLOCKS ? MIN_BLOCKS : want));
    volume_min_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(in_dev, numel, (float *)workspace_dev);
    int rc = check_launch("volume_min_partial_kernel");
    if (rc) return rc;
    volume_min_final_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>((const float *)workspace_dev, blocks, out_dev);
    return check_launch("volume_min_final_kernel");
}
