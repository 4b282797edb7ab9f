// Internal interface between the MIND-SSC C-ABI entry point (mind_ssc.cu) and its kernels.
#pragma once
#include "common.cuh"

namespace dgtta {

// dg_tta/mind.py:104-136.  The six neighbours, indexed 0..5 = D-,D+,H-,H+,W-,W+; channel c is
// nb[P1[c]] - nb[P2[c]] (verified against tests/golden/mind_shift_table.npz).
enum { NB_DM = 0, NB_DP = 1, NB_HM = 2, NB_HP = 3, NB_WM = 4, NB_WP = 5 };
__host__ __device__ constexpr int mind_p1(int c)
{
    constexpr int t[12] = {NB_WM, NB_HM, NB_HM, NB_WP, NB_WP, NB_DP, NB_DP, NB_DP, NB_HP, NB_HP, NB_HP, NB_HP};
    return t[c];
}
__host__ __device__ constexpr int mind_p2(int c)
{
    constexpr int t[12] = {NB_DM, NB_DM, NB_WM, NB_DM, NB_HM, NB_WM, NB_HM, NB_WP, NB_DM, NB_WM, NB_WP, NB_DP};
    return t[c];
}

constexpr int MIND_TH = 16;  // patch rows per CTA
constexpr int MIND_TW = 32;  // patch columns per CTA
constexpr int MIND_THREADS = MIND_TH * MIND_TW;

struct MindArgs {
    const float *img;
    float *out;
    const float *noise;
    const float *in_scale;
    void *workspace;
    size_t workspace_bytes;
    int B, D, H, W;
    int delta;
    int ntaps;
    int noise_mode;
    float rw;
    float taps[9];
};

// fast path: R == 2 (5 taps), delta in {1,2,3}, noise NONE/TENSOR.  Returns DGTTA_EUNSUPPORTED otherwise.
bool mind_fast_supported(const MindArgs &a);
size_t mind_fast_workspace_bytes(int B, int D, int H, int W);
int mind_fast_launch(const MindArgs &a, cudaStream_t stream);

// general path: any tap radius 1..4, any delta
size_t mind_general_workspace_bytes(int B, int D, int H, int W);
int mind_general_launch(const MindArgs &a, cudaStream_t stream);

}  // namespace dgtta
