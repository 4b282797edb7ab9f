// MIND-SSC descriptor — tuned sm_100a kernel for the reference's configuration: 5-tap Gaussian
// (sigma = 1, dg_tta/mind.py:31) and dilation delta in {1,2,3}.  Math: see mind_ssc.cu.
//
// One CTA (512 threads, one per SM) owns a 16 x 32 patch of the H-W plane and marches along D in
// batches of PB = 4 planes.  All arithmetic is packed fp32x2 (FFMA2/FMUL2/FADD2).  Per batch:
//   T  ring of raw image planes in shared memory, only PB new planes per batch.
//      LDG modes : cp.async, tile[slot][20+2d][44] holds I at clamped coordinates (the replicate padding of
//                  mind.py:137 is folded into the addresses).
//      TMA mode  : one 48-column box per plane (UTMALDG.3D, zeros outside the volume); S1 clamps the rows it
//                  reads and patches the W neighbours of the volume's first / last quad instead.
//   S1 tasks (plane, halo row, 4 columns) — 720 (36-column tile) or 800 (TMA mode, aligned 40-column tile): the
//      task loads the 6-neighbourhood of its 4 positions (LDS.128), adds the noise and writes the 12 squared edge
//      channels -> ws[plane][c][row][col] (STS.128).  Noise: LDG modes read it from global memory (L2-prefetched);
//      TMA mode finds it already in ws — one 40 x 21 x 12-channel box per plane (UTMALDG.4D) lands where E^2
//      goes and is squared in place; the four ws planes are a ring refilled (plane z+4) as soon as stage C has the
//      plane's values in registers (full / empty mbarriers).
//   S2 tasks (plane, channel, 4-column strip), one per thread: H smoothing with a sliding 5-row register window,
//      in place (LDS.128 / STS.128); conflict-free in both layouts (channel pitch 185 resp. 210 16-byte chunks).
//   C  warp = patch row; thread = (4 consecutive voxels) x (3 channels); lanes = 8 column groups x 4
//      channel groups.  Per plane: the 8 values of the thread's W windows per channel (2 LDS.128, or
//      LDS.64 + LDS.128 + LDS.64 in the 40-column tile) give 4 W-smoothed values; the D smoothing runs
//      on a 4-slot register window (slots are compile-time because PB == taps - 1); min / mean over the 12
//      channels by two xor-shuffles; exp; one STG.128 per channel (full 128-byte lines per warp).
// Three __syncthreads per 4 planes; HBM traffic is 4 B/voxel in, 48 B/voxel of noise when it is streamed in (the
// x1.6 box over-read is served by L2 as long as neighbouring CTAs march in step: ncu dram__bytes_read = 0.737 GB
// = the algorithmic 0.736 GB on 2 x 192^3) and 48 B/voxel out.
//
// Volume faces (TMA mode).  The replicate padding of E^2 (mind.py:22) costs nothing extra: S1 computes only quads
// inside the volume, S2 reads the clamped ROW (rows above the volume re-read row rT, rows below keep the registers of
// row rB) and stage C replaces the two COLUMNS left of w = 0 / right of w = W-1 by the border column.  The passes act
// on different axes, so padding one axis after smoothing another is bit-identical to padding E^2 first.  (Round 1
// let the S1 task owning a border position write the replicated copies: patches on a face then ran 1.25x, corner
// patches 1.55x longer than interior ones — the kernel's duration was the corner patches' — and their lag behind
// the neighbours defeated the L2 sharing of the halo boxes: 1.07 GB read.)
//
// Global clamp (mind.py:158-160): pass 1 assumes it inactive and records {sum v, min positive v, max v}
// per (CTA, batch).  Pass 2 (mind_fast_fix_kernel) reduces them to mean_all(v) in every CTA and recomputes exactly the
// 4-plane units whose range leaves [0.001*mean, 1000*mean] with the clamp.  Two launches per call.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <atomic>
#include <type_traits>

#include "mind_internal.cuh"

namespace dgtta {

namespace fast {

// internal noise mode: DGTTA_NOISE_TENSOR whose field is staged by TMA (cp.async.bulk.tensor) straight into the
// E^2 planes of shared memory, one 40 x 21 x 12-channel box per plane, and squared in place by S1.  TMA wants the
// innermost start coordinate 16-byte aligned (measured: a start of w0-2 floats raises an illegal-instruction trap,
// tools/microbench/tma_probe.cu), so this mode widens the halo tile to the aligned columns [w0-4, w0+36).  Needs
// W % 4 == 0 and a 16-byte aligned tensor; everything else keeps the LDG path.
constexpr int NOISE_TMA = 3;

constexpr int R = 2, NT = 5, PB = 4, NWIN = NT - 1;   // PB == NWIN keeps the D-window slots compile-time
// Patch rows per CTA.  16 rows = one 512-thread CTA per SM; 8 rows = two independent 256-thread CTAs per SM whose
// shared-memory-bound phases (S1, S2) overlap the other CTA's FP32-bound phase (C) at the price of more halo rows.
#ifndef DGTTA_FAST_TH
#define DGTTA_FAST_TH 16
#endif
constexpr int TH = DGTTA_FAST_TH, TW = MIND_TW;
constexpr int CTAS_PER_SM = TH == 16 ? 1 : 2;
constexpr int EH = TH + 2 * R;             // halo rows: 20 / 12
constexpr int C_WARPS = TH;                // one warp per patch row in stage C
constexpr int NTHREADS = TH * 32;          // 512 / 256

// Halo-tile geometry per noise mode.  LDG modes: 36 halo columns [w0-2, w0+34), ws channel pitch 740 words = 185
// 16-byte chunks == 1 (mod 8) so that consecutive S2 tasks (9 strips per channel) stay conflict-free.  TMA mode: 40
// columns [w0-4, w0+36) (the outer two on each side are computed but never read), dense box layout (800 words).
template <int NOISE> struct Mode {
    static constexpr bool TMA = NOISE == NOISE_TMA;
    static constexpr int CO = TMA ? 4 : R;             // halo column 0 is w0 - CO
    static constexpr int EW = TW + 2 * CO;        // halo columns: 40 / 36
    static constexpr int NQUAD = EW / 4;               // position quads per halo row: 10 / 9
    static constexpr int TWD = EW + 8;                 // image tile row pitch: 4 pad + EW + 4 pad words
    static constexpr int WSP = EW;                     // ws row pitch
    // TMA: the box carries one extra row so that the channel pitch is 21 * 10 = 210 chunks == 2 (mod 8): with 10 strips
    // per channel, S2 task t = 10 c + s then hits chunk == t (mod 8) -> conflict-free (a dense 20-row box gives 200 == 0)
    static constexpr int BOX_ROWS = TMA ? EH + 1 : EH;
    static constexpr int WS_CH = TMA ? BOX_ROWS * WSP : EH * WSP + 20;
    static constexpr int WS_PLANE = 12 * WS_CH;
    // ws plane ring = the PB planes of a batch.  A spare fifth plane (noise of the next batch's first plane in flight
    // during all of C) was measured slower (0.608 vs 0.589 ms): the ring arithmetic costs registers in stage C, and the
    // last plane's box — issued at the end of C — is only needed by S1's second task round anyway.
    static constexpr int NSW = PB;
    static constexpr int S1_TASKS = PB * EH * NQUAD;   // 800 / 720
    static constexpr int S2_TASKS = PB * 12 * NQUAD;   // 480 / 432 column-strip tasks
};

template <int DELTA, int NOISE>
struct Geom {
    using M = Mode<NOISE>;
    static constexpr int TR = EH + 2 * DELTA;        // tile rows
    static constexpr int TCOLS = M::EW + 2 * DELTA;  // loaded tile columns
    static constexpr int NSLOT = PB + 2 * DELTA;     // ring of image planes
    static constexpr int TILE = TR * M::TWD;
    static constexpr int CELLS = TR * TCOLS;
    static constexpr int NCELL = (CELLS + NTHREADS - 1) / NTHREADS;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(M::NSW * M::WS_PLANE + NSLOT * TILE);
};

struct Params {
    alignas(64) CUtensorMap img_map;     // NOISE_TMA: 3-D view (W, H, B*D) of the image, box 48 x tile rows x 1 (zero fill outside)
    alignas(64) CUtensorMap noise_map;   // NOISE_TMA: 4-D view (W, H, D, B*12) of the noise tensor, box 40 x 21 x 1 x 12
    const float *img;
    float *out;
    const float *noise;
    const float *in_scale;
    float4 *stats;        // [ncta * nbatch] {sum v, min positive v, max v, -}
    const int *fix_hdr;   // {count} then list of unit ids (FIX pass)
    const float *fix_lohi;
    int B, D, H, W;
    int nTH, nTW, nCD, chunkD, nbatch;
    float rw;
    float taps[NT];
};

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- mbarrier / TMA (sm_90+ PTX; SASS SYNCS.* / UTMALDG)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(float *dst_smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Hides a value's provenance from the optimiser: per-thread constants derived from opaque(tid) inside the batch loop
// are recomputed each batch (a dozen integer instructions) instead of living in registers across stage C, which
// runs at the 128-register cap.
__device__ __forceinline__ int opaque(int x)
{
    asm volatile("" : "+r"(x));
    return x;
}

__device__ __forceinline__ void tma_load_3d(float *dst_smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void ld4(const float *p, float *v)
{
    const float4 f = *reinterpret_cast<const float4 *>(p);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2): two lanes per issue slot.  The FP32 pipe still
// retires 128 lane-ops/clk/SM (measured, tools/microbench/ffma2.cu), but the freed issue slots carry the
// LDS/STS/SHFL/MUFU traffic of the stencil.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fsub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ void ld4p(const float *p, u64 &a, u64 &b)
{
    const float4 f = *reinterpret_cast<const float4 *>(p);
    a = pk(f.x, f.y); b = pk(f.z, f.w);
}
__device__ __forceinline__ void st4p(float *p, u64 a, u64 b)
{
    float4 f;
    unpk(a, f.x, f.y); unpk(b, f.z, f.w);
    *reinterpret_cast<float4 *>(p) = f;
}

// S1 task index within a plane -> (halo row, position quad).  LDG modes: row-major, 9 quads per row.  TMA mode (10 quads
// per row, 20 rows): groups of 40 tasks = 4 rows; the first 32 are quads 0-7 of the four rows, the last 8 are quads 8-9.
// A 128-bit shared-memory access is served 8 lanes at a time; with this order every 8-lane group reads one row's
// consecutive chunks (conflict-free in the ws planes AND the 48-column image tiles), except the quad-8/9 group, which is
// conflict-free in ws (row pitch 10 chunks) and 2-way in the image tiles.  Row-major order made most image-tile groups
// straddle two rows 12 chunks apart (== 4 mod 8): 2-way conflicts on 7 of the 19 loads of every task.
template <bool TMA_MODE>
__device__ __forceinline__ void s1_row_quad(int rem, int &r, int &q)
{
    if (TMA_MODE) {
        static_assert(EH % 4 == 0 || !TMA_MODE, "row groups of four");
        const int g = rem / 40, u = rem - g * 40;
        if (u < 32) { r = 4 * g + (u >> 3); q = u & 7; }
        else { r = 4 * g + ((u - 32) >> 1); q = 8 + (u & 1); }
    } else {
        constexpr int NQ = Mode<DGTTA_NOISE_NONE>::NQUAD;
        r = rem / NQ; q = rem - r * NQ;
    }
}

// March one CTA over output planes [d0, d1) of patch (h0, w0) of sample b.
template <int DELTA, int NOISE, bool FIX>
__device__ __forceinline__ void process(const Params &P, float *smem, float (*red)[C_WARPS], uint64_t *full_bar,
                                        uint64_t *empty_bar, uint64_t *img_bar, int b, int h0, int w0, int d0, int d1, float lo, float hi,
                                        float4 *stats)
{
    using G = Geom<DELTA, NOISE>;
    using M = Mode<NOISE>;
    constexpr bool TMA = M::TMA;
    constexpr int CO = M::CO, NQUAD = M::NQUAD, TWD = M::TWD, WSP = M::WSP;
    constexpr int WS_CH = M::WS_CH, WS_PLANE = M::WS_PLANE, NSW = M::NSW, S1_TASKS = M::S1_TASKS, S2_TASKS = M::S2_TASKS;
    float *ws = smem;
    float *tiles = smem + NSW * WS_PLANE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = P.D, H = P.H, W = P.W, HW = H * W;
    const float *img = P.img + (size_t)b * D * HW;

    u64 G2[NT];   // taps broadcast to both halves (uniform -> FFMA2 takes them as UR.F32 operands)
#pragma unroll
    for (int t = 0; t < NT; ++t) G2[t] = pk(P.taps[t], P.taps[t]);

    // ---- T: image tile cells of this thread (plane-invariant, recomputed per call: see opaque())
    int loaded_hi;  // highest real plane resident in the ring
    auto load_until = [&](int need_hi) {
        need_hi = min(need_hi, D - 1);
        if (TMA) {
            // one box per image plane (48 columns from the 16-byte aligned w0-8, all tile rows; zeros outside the
            // volume: S1 clamps the rows it reads and patches the W neighbours of the volume's first / last quad), all
            // planes of the batch on one mbarrier.  Every batch arms the barrier, also with zero new planes.
            if (tid == 0) {
                const int np = max(need_hi - loaded_hi, 0);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (np == 0) mbar_arrive(img_bar);
                else mbar_expect_tx(img_bar, (unsigned)(np * G::TILE * sizeof(float)));
                for (int p = loaded_hi + 1; p <= need_hi; ++p)
                    tma_load_3d(tiles + (p % G::NSLOT) * G::TILE, &P.img_map, img_bar, w0 - CO - 4, h0 - R - DELTA, b * D + p);
            }
            loaded_hi = max(loaded_hi, need_hi);
            return;
        }
        if (loaded_hi >= need_hi) { cp_async_commit(); return; }
        const int t0 = opaque(tid);
        int cell_s[G::NCELL], cell_g[G::NCELL];
#pragma unroll
        for (int k = 0; k < G::NCELL; ++k) {
            const int i = t0 + k * NTHREADS;
            const int rr = i / G::TCOLS, cc = i - rr * G::TCOLS;
            cell_s[k] = i < G::CELLS ? rr * TWD + (4 - DELTA) + cc : -1;
            cell_g[k] = clampi(h0 - R - DELTA + rr, 0, H - 1) * W + clampi(w0 - CO - DELTA + cc, 0, W - 1);
        }
        for (int p = loaded_hi + 1; p <= need_hi; ++p) {
            float *dst = tiles + (p % G::NSLOT) * G::TILE;
            const float *src = img + (size_t)p * HW;
#pragma unroll
            for (int k = 0; k < G::NCELL; ++k)
                if (cell_s[k] >= 0) cp_async4(dst + cell_s[k], src + cell_g[k]);
        }
        loaded_hi = need_hi;
        cp_async_commit();
    };

    // halo rows / columns outside the volume replicate E^2 of the clamped position (mind.py:22)
    const int rT = max(0, R - h0), rB = min(EH - 1, H - 1 - h0 + R);
    const int eR = W - 1 - w0 + CO;              // halo column index of w = W-1
    const bool scaled = P.in_scale != nullptr;
    const u64 RW = pk(P.rw, P.rw);
    const bool noise_vec = (NOISE != DGTTA_NOISE_NONE) && ((W & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.noise) & 7) == 0);

    // ---- C bookkeeping: warp = patch row, lane = (4-column group wl, channel group cg)
    const bool c_active = warp < C_WARPS;
    const int wl = lane & 7, cg = lane >> 3;
    const int c_off = (3 * cg) * WS_CH + warp * WSP + 4 * wl;   // halo columns 4wl .. : the thread's voxels are 4wl+CO-R ..
    const int vh = h0 + warp, vw = w0 + 4 * wl;
    const bool row_ok = c_active && vh < H;
    auto valid = [&](int k) { return row_ok && vw + k < W; };
    // per-channel output pointers of the thread's 4-voxel run, advanced plane by plane
    // running output pointer of the thread's 4-voxel run (channel 3*cg), advanced by one plane per emit
    float *op = P.out + (((size_t)b * 12 + 3 * cg) * D + d0) * HW + (size_t)vh * W + vw;
    const size_t ch_stride = (size_t)D * HW;
    const bool vec_store = ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0) && ((W & 3) == 0) && valid(3);
    const bool stat_lane = cg == 0;

    u64 win[3][2][NWIN];   // the last 4 W-smoothed planes of the thread's 4 voxels x 3 channels
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < NWIN; ++t) win[a][j][t] = 0ull;

    const int z_begin = d0 - R, z_end = d1 + R;
    const int nb = (z_end - z_begin + PB - 1) / PB;

    // ---- NOISE_TMA: marched plane g (= z_begin + g) lives in ws slot g % NSW.  One thread arms full_bar[slot] and
    // issues the box copy; S1 waits on the barrier, adds the noise it finds at its own E^2 address and overwrites it.
    // A slot is refilled (plane g + NSW) by the last warp that finishes reading plane g in stage C.
    auto issue_noise = [&](int s, int z) {   // marched plane z -> slot s
        if (z >= z_end) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads/writes of the slot before the async write
        mbar_expect_tx(&full_bar[s], (unsigned)(WS_PLANE * sizeof(float)));
        tma_load_4d(ws + s * WS_PLANE, &P.noise_map, &full_bar[s], w0 - CO, h0 - R, clampi(z, 0, D - 1), b * 12);
    };
    // slot of plane pz of the current batch = slot0 + pz (mod NSW); bit s of fill_parity = parity of the fill S1 waits for
    int slot0 = 0;
    unsigned fill_parity = 0;
    auto slot_of = [&](int pz) { int sl = slot0 + pz; return TMA ? (sl >= NSW ? sl - NSW : sl) : pz; };
    if (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSW; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], C_WARPS); }
            mbar_init(img_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int g = 0; g < NSW; ++g) issue_noise(g, z_begin + g);
        }
        // the first __syncthreads of the batch loop publishes the barriers before anyone waits on them
    }

    // first batch: everything it needs
    loaded_hi = max(clampi(z_begin, 0, D - 1) - DELTA, 0) - 1;
    load_until(clampi(z_begin + PB - 1, 0, D - 1) + DELTA);

    for (int n = 0; n < nb; ++n) {
        const int zb = z_begin + n * PB;
        if (!TMA) cp_async_wait_all();
        __syncthreads();   // tiles of this batch visible; every thread finished C of the previous batch
        if (!FIX && n > 0 && warp == 0) {
            // statistics of the previous batch (warp partials were parked in `red` before the barrier)
            float s = lane < C_WARPS ? red[0][lane] : 0.f;
            float mnv = lane < C_WARPS ? red[1][lane] : __int_as_float(0x7f800000);
            float mxv = lane < C_WARPS ? red[2][lane] : 0.f;
            s = warp_sum(s); mnv = warp_min(mnv); mxv = warp_max(mxv);
            if (lane == 0) stats[n - 1] = make_float4(s, mnv, mxv, 0.f);
        }

        // ================= S1: squared edges of 4 positions x 12 channels per task
#pragma unroll 1
        for (int t = tid; t < S1_TASKS; t += NTHREADS) {
            const int pz = t / (EH * NQUAD);
            const int rem = t - pz * (EH * NQUAD);
            int r, q;
            s1_row_quad<TMA>(rem, r, q);
            if (zb + pz >= z_end) continue;
            if (TMA && (r < rT || r > rB || w0 - CO + 4 * q < 0 || w0 - CO + 4 * q >= W)) continue;   // not an owner (see below)
            const int zc = clampi(zb + pz, 0, D - 1);
            const int rc = clampi(r, rT, rB);
            const int toff = (rc + DELTA) * TWD + 4 + 4 * q;   // centre row, first position of the quad
            const float *tc = tiles + (zc % G::NSLOT) * G::TILE + toff;
            const float *tm = tiles + (clampi(zc - DELTA, 0, D - 1) % G::NSLOT) * G::TILE + toff;
            const float *tp = tiles + (clampi(zc + DELTA, 0, D - 1) % G::NSLOT) * G::TILE + toff;
            u64 nbv[6][2];   // D-, D+, H-, H+, W-, W+ as (k0,k1),(k2,k3)
            const int gw0 = w0 - CO + 4 * q;
            int hm_off = -DELTA * TWD, hp_off = DELTA * TWD;
            if (TMA) {
                // image tiles arrive by TMA with zeros outside the volume: read the clamped row instead (replicate
                // padding, mind.py:137); tile row i is h = h0 - R - DELTA + i
                mbar_wait(img_bar, (unsigned)n & 1u);
                const int row_lo = max(0, R + DELTA - h0), row_hi = min(G::TR - 1, H - 1 - h0 + R + DELTA);
                hm_off = (max(rc, row_lo) - (rc + DELTA)) * TWD;
                hp_off = (min(rc + 2 * DELTA, row_hi) - (rc + DELTA)) * TWD;
            }
            ld4p(tm, nbv[NB_DM][0], nbv[NB_DM][1]);
            ld4p(tp, nbv[NB_DP][0], nbv[NB_DP][1]);
            ld4p(tc + hm_off, nbv[NB_HM][0], nbv[NB_HM][1]);
            ld4p(tc + hp_off, nbv[NB_HP][0], nbv[NB_HP][1]);
            float wr[12];      // centre row, positions -4 .. 7 relative to the quad
            ld4(tc - 4, &wr[0]); ld4(tc, &wr[4]); ld4(tc + 4, &wr[8]);
            float wm[4], wp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { wm[k] = wr[4 + k - DELTA]; wp[k] = wr[4 + k + DELTA]; }
            if (TMA) {
                // same for the W neighbours of the volume's first and last quad (W % 4 == 0: quads are aligned)
                if (gw0 == 0) {
#pragma unroll
                    for (int k = 0; k < DELTA; ++k) wm[k] = wr[4];
                }
                if (gw0 + 4 == W) {
#pragma unroll
                    for (int k = 4 - DELTA; k < 4; ++k) wp[k] = wr[7];
                }
            }
            if (!TMA && w0 == 0 && q == 0) {
                // columns w < 0 take E of w = 0: only the W+ neighbour (I at w = delta) differs from the clamped loads
#pragma unroll
                for (int k = 0; k < CO; ++k) wp[k] = wr[4 + CO + DELTA];
            }
            if (!TMA && 4 * q + 3 > eR) {
                // columns beyond w = W-1 take E of w = W-1: only the W- neighbour differs
                const float wm_fix = tiles[(zc % G::NSLOT) * G::TILE + (rc + DELTA) * TWD + 4 + eR - DELTA];
#pragma unroll
                for (int k = 0; k < 4; ++k) if (4 * q + k > eR) wm[k] = wm_fix;
            }
            nbv[NB_WM][0] = pk(wm[0], wm[1]); nbv[NB_WM][1] = pk(wm[2], wm[3]);
            nbv[NB_WP][0] = pk(wp[0], wp[1]); nbv[NB_WP][1] = pk(wp[2], wp[3]);
            if (scaled) {
                const float sa = __ldg(P.in_scale + 2 * b), sc = __ldg(P.in_scale + 2 * b + 1);
                const u64 SA = pk(sa, sa), SC = pk(sc, sc);
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int j = 0; j < 2; ++j) nbv[a][j] = fmul2(fmul2(nbv[a][j], SA), SC);
            }
            float *wsp = ws + slot_of(pz) * WS_PLANE + r * WSP + 4 * q;
            if (TMA) {
                // Only quads inside the volume are computed (W % 4 == 0: a quad is in or out as a whole).  Halo positions
                // outside replicate E^2 of the clamped position (mind.py:22): S2 reads the clamped ROW and stage C the
                // clamped COLUMN instead — the smoothing passes act on different axes, so padding one axis after
                // smoothing another gives bit-identical values, and no task writes replicated copies.
                const int sl = slot_of(pz);
                mbar_wait(&full_bar[sl], (fill_parity >> sl) & 1u);   // also orders our E^2 stores after the box write
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    u64 n0, n1;
                    ld4p(wsp + c * WS_CH, n0, n1);
                    u64 e0 = fsub2(nbv[mind_p1(c)][0], nbv[mind_p2(c)][0]);
                    u64 e1 = fsub2(nbv[mind_p1(c)][1], nbv[mind_p2(c)][1]);
                    e0 = fadd2(e0, fmul2(RW, n0));   // product and sum rounded separately, like the reference
                    e1 = fadd2(e1, fmul2(RW, n1));
                    st4p(wsp + c * WS_CH, fmul2(e0, e0), fmul2(e1, e1));
                }
            } else if (NOISE != DGTTA_NOISE_NONE) {
                // noise of the (clamped) positions: mind.py:150-152 adds rw*N to the edge before squaring; halo
                // positions outside the volume reuse the sample of the clamped position (they replicate E^2)
                const float *nz = P.noise + ((size_t)b * 12 * D + zc) * HW + (size_t)clampi(h0 - R + r, 0, H - 1) * W;
                const bool interior = gw0 >= 0 && gw0 + 3 < W && noise_vec;   // uniform per task, 8-byte aligned pairs
                int ngw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ngw[k] = clampi(gw0 + k, 0, W - 1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    u64 n0[6], n1[6];   // six channels in flight at a time
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) {
                        const float *np = nz + (size_t)(6 * half + cc) * D * HW;
                        if (interior) {
                            const float2 a = __ldg(reinterpret_cast<const float2 *>(np + gw0));
                            const float2 bq = __ldg(reinterpret_cast<const float2 *>(np + gw0 + 2));
                            n0[cc] = pk(a.x, a.y); n1[cc] = pk(bq.x, bq.y);
                        } else {
                            n0[cc] = pk(__ldg(np + ngw[0]), __ldg(np + ngw[1]));
                            n1[cc] = pk(__ldg(np + ngw[2]), __ldg(np + ngw[3]));
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) {
                        const int c = 6 * half + cc;
                        u64 e0 = fsub2(nbv[mind_p1(c)][0], nbv[mind_p2(c)][0]);
                        u64 e1 = fsub2(nbv[mind_p1(c)][1], nbv[mind_p2(c)][1]);
                        e0 = fadd2(e0, fmul2(RW, n0[cc]));   // product and sum rounded separately, like the reference
                        e1 = fadd2(e1, fmul2(RW, n1[cc]));
                        st4p(wsp + c * WS_CH, fmul2(e0, e0), fmul2(e1, e1));
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const u64 e0 = fsub2(nbv[mind_p1(c)][0], nbv[mind_p2(c)][0]);
                    const u64 e1 = fsub2(nbv[mind_p1(c)][1], nbv[mind_p2(c)][1]);
                    st4p(wsp + c * WS_CH, fmul2(e0, e0), fmul2(e1, e1));
                }
            }
        }
        __syncthreads();

        // ================= S2: H smoothing of a 4-column strip, in place, sliding 5-row register window
        // one (plane, channel, column quad) task per thread, exactly one round
#pragma unroll 1
        for (int s2_t = opaque(tid); s2_t < S2_TASKS; s2_t += NTHREADS) {
        const int s2_pc = s2_t / NQUAD;                  // pz * 12 + c
        const int s2_pz = s2_pc / 12;
        const int s2_gw = w0 - CO + 4 * (s2_t - s2_pc * NQUAD);   // first column of the strip
        if (zb + s2_pz < z_end && !(TMA && (s2_gw < 0 || s2_gw >= W))) {   // (strips outside the volume: never read)
            float *col = ws + slot_of(s2_pz) * WS_PLANE + (s2_pc - s2_pz * 12) * WS_CH + 4 * (s2_t - s2_pc * NQUAD);
            u64 acc[NT][2];   // partial sums of the 5 output rows that input row r contributes to
            u64 x0 = 0ull, x1 = 0ull;
#pragma unroll
            for (int r = 0; r < EH; ++r) {
                if (TMA) {
                    // rows outside the volume replicate the first / last row inside (mind.py:22): rows < rT read row rT
                    // (rT <= R), rows > rB keep the registers of row rB
                    if (r < R) ld4p(col + max(r, rT) * WSP, x0, x1);
                    else if (r <= rB) ld4p(col + r * WSP, x0, x1);
                } else {
                    ld4p(col + r * WSP, x0, x1);
                }
                // out_o = sum_t g_t x[o+t]: row r feeds o = r-t with tap t; ascending t per output = ascending rows
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const int o = r - t;
                    if (o < 0 || o >= TH) continue;
                    if (t == 0) { acc[o % NT][0] = fmul2(x0, G2[0]); acc[o % NT][1] = fmul2(x1, G2[0]); }
                    else { acc[o % NT][0] = ffma2(x0, G2[t], acc[o % NT][0]); acc[o % NT][1] = ffma2(x1, G2[t], acc[o % NT][1]); }
                }
                const int done = r - (NT - 1);
                if (done >= 0) st4p(col + done * WSP, acc[done % NT][0], acc[done % NT][1]);   // row `done` is dead as an input
            }
        }
        }
        __syncthreads();   // ws complete; tiles no longer read in this batch

        // prefetch the image planes of the next batch while C runs
        if (n + 1 < nb) {
            load_until(clampi(zb + 2 * PB - 1, 0, D - 1) + DELTA);
            if (NOISE == DGTTA_NOISE_TENSOR) {   // (NOISE_TMA: the boxes are already in flight)
                // pull the next batch's noise rows (144 B each) into L2 so that S1's loads do not pay DRAM latency
                for (int i = tid; i < 12 * PB * EH; i += NTHREADS) {
                    const int c = i / (PB * EH), rem = i - c * (PB * EH);
                    const int pz = rem / EH, r = rem - pz * EH;
                    const int zn = zb + PB + pz;
                    if (zn >= z_end) continue;
                    const float *np = P.noise + (((size_t)b * 12 + c) * D + clampi(zn, 0, D - 1)) * HW +
                                      (size_t)clampi(h0 - R + r, 0, H - 1) * W + max(w0 - R, 0);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(np));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(np + min(32, W - 1 - max(w0 - R, 0))));
                }
            }
        }

        // ================= C: W smoothing, D window, MIND normalisation, store
        float st_sum = 0.f, st_max = 0.f;
        unsigned st_minu = 0xffffffffu;   // (bits of the smallest positive v) - 1
        if (c_active) {
#pragma unroll
            for (int ph = 0; ph < PB; ++ph) {
                const int z = zb + ph;
                {
                    // No branch around a plane: planes past z_end (tail of the last batch) read stale but finite ws
                    // data and are never emitted, so the four plane steps form one straight-line block that the
                    // scheduler can overlap (loads / W pass of plane p+1 under the shuffle+MUFU chain of plane p).
                    const float *wsp = ws + slot_of(ph) * WS_PLANE + c_off;
                    const bool emit = (z - R) >= d0 && z < z_end;   // uniform
                    u64 m[3][2];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        float v[8];   // the 5-tap window of voxel k starts at v[k] = halo column 4wl + k + CO - R
                        if (TMA) {
                            const float2 f0 = *reinterpret_cast<const float2 *>(wsp + a * WS_CH + 2);
                            const float2 f1 = *reinterpret_cast<const float2 *>(wsp + a * WS_CH + 8);
                            v[0] = f0.x; v[1] = f0.y; v[6] = f1.x; v[7] = f1.y;
                            ld4(wsp + a * WS_CH + 4, &v[2]);
                            // columns outside the volume replicate the first / last column inside (mind.py:22).  (Two code
                            // copies selected once per CTA — selects only in patches on a W face — ran slower: the doubled
                            // instruction footprint showed up as run-to-run variance on every CTA.)
                            if (vw == 0) { v[0] = v[2]; v[1] = v[2]; }
                            if (vw + 4 == W) { v[6] = v[5]; v[7] = v[5]; }
                        } else {
                            ld4(wsp + a * WS_CH, &v[0]);
                            ld4(wsp + a * WS_CH + 4, &v[4]);
                        }
                        // o_k = g2 v[k+2] + g1 (v[k+1] + v[k+3]) + g0 (v[k] + v[k+4]),  k = 0..3, two lanes at a time
                        const u64 s0a = fadd2(pk(v[0], v[1]), pk(v[4], v[5]));
                        const u64 s0b = fadd2(pk(v[2], v[3]), pk(v[6], v[7]));
                        const u64 s1a = pk(v[1] + v[3], v[2] + v[4]);
                        const u64 s1b = pk(v[3] + v[5], v[4] + v[6]);
                        u64 oa = fmul2(pk(v[2], v[3]), G2[2]);
                        u64 ob = fmul2(pk(v[4], v[5]), G2[2]);
                        oa = ffma2(s1a, G2[1], oa); ob = ffma2(s1b, G2[1], ob);
                        oa = ffma2(s0a, G2[0], oa); ob = ffma2(s0b, G2[0], ob);
                        // D smoothing: slots ph, ph+1, ph+2, ph+3 (mod 4) hold planes z-4 .. z-1.  (The scatter form — four
                        // running sums updated in place, no register copies — measured 2-4 % slower.)
                        u64 acc0 = fmul2(win[a][0][ph], G2[0]), acc1 = fmul2(win[a][1][ph], G2[0]);
#pragma unroll
                        for (int t = 1; t < NWIN; ++t) {
                            acc0 = ffma2(win[a][0][(ph + t) % NWIN], G2[t], acc0);
                            acc1 = ffma2(win[a][1][(ph + t) % NWIN], G2[t], acc1);
                        }
                        m[a][0] = ffma2(oa, G2[NT - 1], acc0);
                        m[a][1] = ffma2(ob, G2[NT - 1], acc1);
                        win[a][0][ph] = oa;
                        win[a][1][ph] = ob;
                    }
                    if (TMA) {
                        // the plane's values are in registers: release the slot now, and let thread 0 refill it as soon as
                        // the other warps have got this far too (the last plane's box then leaves ~half a plane step
                        // earlier than at the end of C)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[slot_of(ph)]);
                        if (tid == 0) {
                            const int sl = slot_of(ph);
                            mbar_wait(&empty_bar[sl], (fill_parity >> sl) & 1u);
                            issue_noise(sl, z + NSW);
                        }
                    }
                    float o[3][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float x0[2], x1[2], x2[2];
                        unpk(m[0][j], x0[0], x0[1]); unpk(m[1][j], x1[0], x1[1]); unpk(m[2][j], x2[0], x2[1]);
                        float mn[2], sc2[2], vv[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float t = fminf(fminf(x0[e], x1[e]), x2[e]);
                            t = fminf(t, __shfl_xor_sync(0xffffffffu, t, 8));
                            t = fminf(t, __shfl_xor_sync(0xffffffffu, t, 16));
                            mn[e] = t;
                        }
                        const u64 MN = pk(mn[0], mn[1]);
                        const u64 m0 = fsub2(m[0][j], MN), m1 = fsub2(m[1][j], MN), m2 = fsub2(m[2][j], MN);   // mind.py:156
                        const u64 S = fadd2(fadd2(m0, m1), m2);
                        float s[2];
                        unpk(S, s[0], s[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float t = s[e];
                            t += __shfl_xor_sync(0xffffffffu, t, 8);
                            t += __shfl_xor_sync(0xffffffffu, t, 16);
                            float v = t * (1.f / 12.f);                                   // mind.py:157
                            if (FIX) {
                                v = fminf(fmaxf(v, lo), hi);                               // mind.py:158-160
                                sc2[e] = -1.4426950408889634f * rcp_approx(v);
                            } else {
                                vv[e] = v;
                                // v == 0 <=> all m_c == 0: 0 * (-log2e / tiny) = -0 -> exp = 1 = exp(-0/lo) for any lo > 0
                                // (lo == 0 is caught by pass 2, and so is every v below 1e-37: it is < 0.001 mean(v))
                                sc2[e] = -1.4426950408889634f * rcp_approx(fmaxf(v, 1e-37f));
                            }
                        }
                        if (!FIX) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                // statistics of the clamp decision: v >= 0, so bit order == value order, and bits - 1 sends
                                // v == 0 to 0xffffffff, out of the minimum over the positive v
                                const float vm = (stat_lane && emit && valid(2 * j + e)) ? vv[e] : 0.f;
                                st_sum += vm;
                                st_max = fmaxf(st_max, vm);
                                st_minu = min(st_minu, __float_as_uint(vm) - 1u);
                            }
                        }
                        const u64 SCL = pk(sc2[0], sc2[1]);
                        float y0[2], y1[2], y2[2];
                        unpk(fmul2(m0, SCL), y0[0], y0[1]); unpk(fmul2(m1, SCL), y1[0], y1[1]); unpk(fmul2(m2, SCL), y2[0], y2[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {                                      // mind.py:161-162
                            o[0][2 * j + e] = ex2_approx(y0[e]);
                            o[1][2 * j + e] = ex2_approx(y1[e]);
                            o[2][2 * j + e] = ex2_approx(y2[e]);
                        }
                    }
                    if (vec_store) {
                        if (emit) {
                            __stcs(reinterpret_cast<float4 *>(op), make_float4(o[0][0], o[0][1], o[0][2], o[0][3]));
                            __stcs(reinterpret_cast<float4 *>(op + ch_stride), make_float4(o[1][0], o[1][1], o[1][2], o[1][3]));
                            __stcs(reinterpret_cast<float4 *>(op + 2 * ch_stride), make_float4(o[2][0], o[2][1], o[2][2], o[2][3]));
                        }
                    } else if (emit) {
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (valid(k)) __stcs(op + a * ch_stride + k, o[a][k]);
                    }
                    op += emit ? HW : 0;
                }
            }
            if (!FIX) {
                float st_min = st_minu == 0xffffffffu ? __int_as_float(0x7f800000) : __uint_as_float(st_minu + 1u);
                st_sum = warp_sum(st_sum); st_min = warp_min(st_min); st_max = warp_max(st_max);
                if (lane == 0) { red[0][warp] = st_sum; red[1][warp] = st_min; red[2][warp] = st_max; }
            }
        }
        if (TMA) {
            // the batch consumed one fill of each of its PB slots; the next batch starts PB slots further round the ring
            if (NSW > PB) {
                fill_parity ^= ((1u << NSW) - 1u) ^ (1u << slot_of(PB));
                slot0 = slot_of(PB);
            } else {
                fill_parity ^= (1u << NSW) - 1u;
            }
        }
    }
    if (!FIX) {
        __syncthreads();
        if (warp == 0) {
            float s = lane < C_WARPS ? red[0][lane] : 0.f;
            float mnv = lane < C_WARPS ? red[1][lane] : __int_as_float(0x7f800000);
            float mxv = lane < C_WARPS ? red[2][lane] : 0.f;
            s = warp_sum(s); mnv = warp_min(mnv); mxv = warp_max(mxv);
            if (lane == 0) stats[nb - 1] = make_float4(s, mnv, mxv, 0.f);
            // batches this (shorter) chunk never ran: neutral entries
            for (int i = nb + lane; i < P.nbatch; i += 32) stats[i] = make_float4(0.f, __int_as_float(0x7f800000), 0.f, 0.f);
        }
    }
    if (TMA && FIX) {
        // the persistent fix kernel re-arms the barriers for its next unit; every issued box has been consumed by S1
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSW; ++s) { mbar_inval(&full_bar[s]); mbar_inval(&empty_bar[s]); }
            mbar_inval(img_bar);
        }
    }
}

__device__ __forceinline__ void decode_cta(const Params &P, int cta, int &b, int &h0, int &w0, int &d0, int &d1)
{
    const int cd = cta % P.nCD; cta /= P.nCD;
    const int tw = cta % P.nTW; cta /= P.nTW;
    const int th = cta % P.nTH; cta /= P.nTH;
    b = cta;
    h0 = th * TH; w0 = tw * TW;
    d0 = cd * P.chunkD;
    d1 = min(P.D, d0 + P.chunkD);
}


#ifdef DGTTA_CTA_TIMES
// developer probe (-DDGTTA_CTA_TIMES, tools/dbg_cta_times.py): per-CTA {smid, start, end (globaltimer ns), cycles} of pass 1.
// It is what showed that patches on a volume face ran 1.25-1.55x longer than interior ones (round 2).
__device__ unsigned long long g_dbg[1024][4];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <int DELTA, int NOISE>
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) mind_fast_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(128) float smem[];
    __shared__ float red[3][C_WARPS];
    __shared__ uint64_t full_bar[PB + 1];
    __shared__ uint64_t empty_bar[PB + 1];
    __shared__ uint64_t img_bar;
    int b, h0, w0, d0, d1;
    decode_cta(P, blockIdx.x, b, h0, w0, d0, d1);
#ifdef DGTTA_CTA_TIMES
    const unsigned long long dbg_t0 = gtimer();
    const long long dbg_c0 = clock64();
#endif
    process<DELTA, NOISE, false>(P, smem, red, full_bar, empty_bar, &img_bar, b, h0, w0, d0, d1, 0.f, 0.f,
                                 P.stats + (size_t)blockIdx.x * P.nbatch);
#ifdef DGTTA_CTA_TIMES
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_dbg[blockIdx.x][0] = smid; g_dbg[blockIdx.x][1] = dbg_t0; g_dbg[blockIdx.x][2] = gtimer();
        g_dbg[blockIdx.x][3] = (unsigned long long)(clock64() - dbg_c0);
    }
#endif
}

// pass 2 (one launch, no separate reduction kernel): every CTA reduces the per-unit statistics to mean_all(v) -> clamp
// bounds (same code on the same data in every CTA: identical bounds), then recomputes, with the clamp, exactly those
// (CTA, batch) units of pass 1 whose range of v leaves [0.001*mean, 1000*mean].  Unit u belongs to CTA u % gridDim.x.
template <int DELTA, int NOISE>
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) mind_fast_fix_kernel(const __grid_constant__ Params P, int nunits,
                                                                            double inv_count)
{
    extern __shared__ __align__(128) float smem[];
    __shared__ float red[3][C_WARPS];
    __shared__ uint64_t full_bar[PB + 1];
    __shared__ uint64_t empty_bar[PB + 1];
    __shared__ uint64_t img_bar;
    __shared__ double dred[NTHREADS / 32];
    __shared__ float s_lo, s_hi, s_mean;
    __shared__ int s_count;
    __shared__ int s_list[64];
    const int tid = threadIdx.x;
    double s = 0.0;
    for (int i = tid; i < nunits; i += NTHREADS) s += (double)P.stats[i].x;
    s = warp_sum(s);
    if ((tid & 31) == 0) dred[tid >> 5] = s;
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < NTHREADS / 32; ++i) tot += dred[i];
        const float mean = (float)(tot * inv_count);
        s_mean = mean;
        s_lo = mean * 0.001f;   // mind.py:158-160
        s_hi = mean * 1000.f;
    }
    __syncthreads();
    const float lo = s_lo, hi = s_hi, mean = s_mean;
    // this CTA's units, 64 candidates at a time
    for (int base = blockIdx.x; base < nunits; base += 64 * gridDim.x) {
        if (tid < 64) {
            const int i = base + tid * gridDim.x;
            if (i < nunits) {
                const float4 st = P.stats[i];
                // units that emitted nothing carry {0, inf, 0}
                const bool touched = st.z > 0.f || st.y < __int_as_float(0x7f800000) || !(mean > 0.f);
                if (touched && (!(mean > 0.f) || st.z > hi || st.y < lo)) s_list[atomicAdd(&s_count, 1)] = i;
            }
        }
        __syncthreads();
        const int count = s_count;
        for (int u = 0; u < count; ++u) {
            const int unit = s_list[u];
            const int cta = unit / P.nbatch, n = unit - cta * P.nbatch;
            int b, h0, w0, d0, d1;
            decode_cta(P, cta, b, h0, w0, d0, d1);
            // batch n of pass 1 emitted planes d0 + PB*n - 2R .. + PB-1 (clipped to the chunk)
            const int lo_d = max(d0, d0 + n * PB - 2 * R), hi_d = min(d1 - 1, d0 + n * PB - 2 * R + PB - 1);
            if (lo_d <= hi_d) process<DELTA, NOISE, true>(P, smem, red, full_bar, empty_bar, &img_bar, b, h0, w0, lo_d, hi_d + 1, lo, hi, nullptr);
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0) s_count = 0;
        __syncthreads();
    }
}

struct Plan {
    int nTH, nTW, nCD, chunkD, ncta, nbatch;
};

static Plan make_plan(int B, int D, int H, int W)
{
    Plan p;
    p.nTH = (H + TH - 1) / TH;
    p.nTW = (W + TW - 1) / TW;
    const long base = (long)B * p.nTH * p.nTW;
    // One CTA per SM is resident.  Choose the number of D chunks that minimises
    //   waves * (planes marched per CTA, rounded up to whole batches, + fixed per-CTA cost):
    // splitting D fills idle SMs but every chunk re-marches 2R warm-up planes.
    const int sms = sm_count() * CTAS_PER_SM;
    const int max_chunks = (D + 15) / 16;
    long best_cost = -1;
    int best = 1;
    for (int ncd = 1; ncd <= max_chunks; ++ncd) {
        const int chunk = (D + ncd - 1) / ncd;
        const int n = (D + chunk - 1) / chunk;
        const long waves = (base * n + sms - 1) / sms;
        const long cost = waves * (((chunk + 2 * R + PB - 1) / PB) * PB + 3);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ncd; }
    }
    p.chunkD = (D + best - 1) / best;
    p.nCD = (D + p.chunkD - 1) / p.chunkD;
    p.ncta = (int)(base * p.nCD);
    p.nbatch = (p.chunkD + 2 * R + PB - 1) / PB;
    return p;
}

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

template <int DELTA, int NOISE>
static int launch(const Params &P0, const Plan &plan, void *workspace, cudaStream_t stream)
{
    using G = Geom<DELTA, NOISE>;
    // the opt-in to > 48 KB of dynamic shared memory is per device: remember it per (instantiation, device)
    static std::atomic<bool> configured_on[64];   // idempotent set-up: a race only repeats it
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::atomic<bool> &configured = configured_on[dev];
    if (!configured) {
        cudaFuncSetAttribute(mind_fast_kernel<DELTA, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        cudaFuncSetAttribute(mind_fast_fix_kernel<DELTA, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        configured = true;
    }
    Params P = P0;
    const int nunits = plan.ncta * plan.nbatch;
    char *base = (char *)workspace;
    P.stats = (float4 *)base;
    int *hdr = (int *)(base + align16((size_t)nunits * sizeof(float4)));
    float *lohi = (float *)((char *)hdr + align16((size_t)(nunits + 1) * sizeof(int)));
    P.fix_hdr = hdr;
    P.fix_lohi = lohi;
    mind_fast_kernel<DELTA, NOISE><<<plan.ncta, NTHREADS, G::SMEM, stream>>>(P);
    int rc = check_launch("mind_fast_kernel");
    if (rc) return rc;
    mind_fast_fix_kernel<DELTA, NOISE><<<sm_count() * CTAS_PER_SM, NTHREADS, G::SMEM, stream>>>(
        P, nunits, 1.0 / ((double)P.B * P.D * P.H * P.W));
    return check_launch("mind_fast_fix_kernel");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    // function-local static: initialised exactly once, thread-safe (C++11)
    static const EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)p;
        cudaGetLastError();
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// developer knob: DGTTA_TMA_PROMO = 0 (none) / 1 (64 B) / 2 (128 B, default) / 3 (256 B) L2 promotion of the TMA boxes
static CUtensorMapL2promotion l2_promotion()
{
    const char *e = getenv("DGTTA_TMA_PROMO");
    const int v = e ? atoi(e) : 2;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}

// 4-D view (W, H, D, B*12) of the [B,12,D,H,W] noise tensor; box = one halo tile of all 12 channels of one plane
static bool make_noise_map(Params &P)
{
    if ((P.W & 3) || (reinterpret_cast<uintptr_t>(P.noise) & 15) || (reinterpret_cast<uintptr_t>(P.out) & 15)) return false;
    if (getenv("DGTTA_MIND_NO_TMA")) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.D, (cuuint64_t)P.B * 12};
    const cuuint64_t strides[3] = {(cuuint64_t)P.W * 4, (cuuint64_t)P.H * P.W * 4, (cuuint64_t)P.D * P.H * P.W * 4};
    const cuuint32_t box[4] = {Mode<NOISE_TMA>::EW, Mode<NOISE_TMA>::BOX_ROWS, 1, 12};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(&P.noise_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(P.noise), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 3-D view (W, H, B*D) of the image; box = one plane of the CTA's tile (48 columns x tile rows), zero fill outside
template <int DELTA>
static bool make_img_map(Params &P)
{
    if (reinterpret_cast<uintptr_t>(P.img) & 15) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.B * P.D};
    const cuuint64_t strides[2] = {(cuuint64_t)P.W * 4, (cuuint64_t)P.H * P.W * 4};
    const cuuint32_t box[3] = {Mode<NOISE_TMA>::TWD, Geom<DELTA, NOISE_TMA>::TR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(&P.img_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(P.img), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int DELTA>
static int launch_noise(Params &P, const Plan &plan, void *workspace, int noise_mode, cudaStream_t stream)
{
    if (noise_mode == DGTTA_NOISE_TENSOR) {
        constexpr int TMA_MAX_DELTA = CTAS_PER_SM == 1 ? 3 : 1;   // shared-memory budget
        if (DELTA <= TMA_MAX_DELTA && make_noise_map(P) && make_img_map<DELTA <= TMA_MAX_DELTA ? DELTA : 1>(P))
            return launch<DELTA <= TMA_MAX_DELTA ? DELTA : 1, NOISE_TMA>(P, plan, workspace, stream);
        return launch<DELTA, DGTTA_NOISE_TENSOR>(P, plan, workspace, stream);
    }
    return launch<DELTA, DGTTA_NOISE_NONE>(P, plan, workspace, stream);
}

}  // namespace fast


#ifdef DGTTA_CTA_TIMES
extern "C" int dgtta_debug_cta_times(unsigned long long *host_out, int n)
{
    return (int)cudaMemcpyFromSymbol(host_out, fast::g_dbg, sizeof(unsigned long long) * 4 * (n < 1024 ? n : 1024));
}
#endif

template <int DELTA, int NOISE>
static void touch_fast()
{
    DGTTA_TOUCH(fast::mind_fast_kernel<DELTA, NOISE>);
    DGTTA_TOUCH(fast::mind_fast_fix_kernel<DELTA, NOISE>);
}

void preload_mind_fast()
{
    touch_fast<1, DGTTA_NOISE_NONE>(); touch_fast<2, DGTTA_NOISE_NONE>(); touch_fast<3, DGTTA_NOISE_NONE>();
    touch_fast<1, DGTTA_NOISE_TENSOR>(); touch_fast<2, DGTTA_NOISE_TENSOR>(); touch_fast<3, DGTTA_NOISE_TENSOR>();
    touch_fast<1, fast::NOISE_TMA>();
    if (fast::CTAS_PER_SM == 1) { touch_fast<fast::CTAS_PER_SM == 1 ? 2 : 1, fast::NOISE_TMA>(); touch_fast<fast::CTAS_PER_SM == 1 ? 3 : 1, fast::NOISE_TMA>(); }
}

bool mind_fast_supported(const MindArgs &a)
{
    return a.ntaps == fast::NT && a.delta >= 1 && a.delta <= 3 &&
           (a.noise_mode == DGTTA_NOISE_NONE || a.noise_mode == DGTTA_NOISE_TENSOR);
}

size_t mind_fast_workspace_bytes(int B, int D, int H, int W)
{
    // bound independent of the SM count: chunks are >= 16 planes (or the whole of D)
    const size_t nTH = (H + fast::TH - 1) / fast::TH, nTW = (W + fast::TW - 1) / fast::TW;
    const size_t max_chunks = (D + 15) / 16;
    // per chunk at most ceil((chunkD + 4)/5) batches with chunkD <= D; bound the product generously
    const size_t units = (size_t)B * nTH * nTW * ((size_t)D / fast::PB + 3 * max_chunks + 2);
    return fast::align16(units * sizeof(float4)) + fast::align16((units + 1) * sizeof(int)) + 64;
}

int mind_fast_launch(const MindArgs &a, cudaStream_t stream)
{
    const fast::Plan plan = fast::make_plan(a.B, a.D, a.H, a.W);
    const size_t nunits = (size_t)plan.ncta * plan.nbatch;
    const size_t need = fast::align16(nunits * sizeof(float4)) + fast::align16((nunits + 1) * sizeof(int)) + 64;
    if (a.workspace_bytes < need) {
        set_error("dgtta_mind_ssc_fwd: workspace too small (%zu < %zu)", a.workspace_bytes, need);
        return DGTTA_EWORKSPACE;
    }
    fast::Params P;
    memset(&P.noise_map, 0, sizeof(P.noise_map));
    memset(&P.img_map, 0, sizeof(P.img_map));
    P.img = a.img; P.out = a.out; P.noise = a.noise; P.in_scale = a.in_scale;
    P.stats = nullptr; P.fix_hdr = nullptr; P.fix_lohi = nullptr;
    P.B = a.B; P.D = a.D; P.H = a.H; P.W = a.W;
    P.nTH = plan.nTH; P.nTW = plan.nTW; P.nCD = plan.nCD; P.chunkD = plan.chunkD; P.nbatch = plan.nbatch;
    P.rw = a.rw;
    for (int i = 0; i < fast::NT; ++i) P.taps[i] = a.taps[i];
    switch (a.delta) {
        case 1: return fast::launch_noise<1>(P, plan, a.workspace, a.noise_mode, stream);
        case 2: return fast::launch_noise<2>(P, plan, a.workspace, a.noise_mode, stream);
        default: return fast::launch_noise<3>(P, plan, a.workspace, a.noise_mode, stream);
    }
}

}  // namespace dgtta
