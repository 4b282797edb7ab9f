// MIND-SSC descriptor - C-ABI entry point (include/dgtta.h).
//
// Replaces MIND3D.forward of the reference (dg_tta/mind.py:142-164; shift table :104-136; Gaussian
// smoothing :5-43).  Math per output voxel p and channel c:
//     E_c(q)   = I(clamp(q + delta*s1_c)) - I(clamp(q + delta*s2_c)) + rw * N_c(q)
//     ssd_c(p) = sum_{ijk} g_i g_j g_k E_c^2(clamp(p + (i,j,k) - R))
//     m_c = ssd_c - min_c ssd_c ;  v = mean_c m_c ;  v <- clamp(v, 0.001*mean_all(v), 1000*mean_all(v))
//     out_c = exp(-m_c / v)
// Two kernels implement it: mind_fast.cu (5 taps, delta <= 3: everything the reference's trainers and
// TTA use) and mind_general.cu (other tap counts / dilations).
#include "mind_internal.cuh"

using namespace dgtta;

namespace dgtta {
int philox_normal_fill(float *out, unsigned long long numel, unsigned long long seed, unsigned long long offset, int sms,
                       int max_threads_per_sm, cudaStream_t stream, const unsigned long long *state_dev = nullptr);   // philox_normal.cu
}

extern "C" size_t dgtta_mind_workspace_bytes(int B, int D, int H, int W)
{
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    const size_t a = mind_fast_workspace_bytes(B, D, H, W), b = mind_general_workspace_bytes(B, D, H, W);
    return a > b ? a : b;
}

extern "C" uint64_t dgtta_mind_philox_offset_increment(int B, int D, int H, int W, int sm_count_, int max_threads_per_sm)
{
    // ATen/native/cuda/DistributionTemplates.h calc_execution_policy: block 256, unroll 4
    const uint64_t numel = (uint64_t)B * 12 * D * H * W;
    const uint64_t block = 256, unroll = 4;
    uint64_t grid = (numel + block - 1) / block;
    const uint64_t cap = (uint64_t)sm_count_ * (max_threads_per_sm / block);
    if (grid > cap) grid = cap;
    return ((numel - 1) / (block * grid * unroll) + 1) * 4;  // curand4 engine calls * 4
}

extern "C" int dgtta_mind_ssc_fwd(const float *img_dev, float *out_dev, const float *in_scale_dev, int B, int D,
                                  int H, int W, int delta, const float *taps_host, int ntaps,
                                  float randn_weighting, int noise_mode, const float *noise_dev,
                                  uint64_t philox_seed, uint64_t philox_offset, void *workspace_dev,
                                  size_t workspace_bytes, dgtta_stream_t stream)
{
    if (!img_dev || !out_dev || !taps_host || !workspace_dev) { set_error("dgtta_mind_ssc_fwd: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || delta < 1 || ntaps < 3 || ntaps > 9 || !(ntaps & 1)) {
        set_error("dgtta_mind_ssc_fwd: bad shape/params B=%d D=%d H=%d W=%d delta=%d ntaps=%d", B, D, H, W, delta, ntaps);
        return DGTTA_EINVAL;
    }
    if ((size_t)B * 12 * D * H * W >= (size_t)1 << 40 || (size_t)H * W >= (size_t)1 << 30) {
        set_error("dgtta_mind_ssc_fwd: volume too large");
        return DGTTA_EINVAL;
    }
    if (noise_mode != DGTTA_NOISE_NONE && noise_mode != DGTTA_NOISE_TENSOR && noise_mode != DGTTA_NOISE_PHILOX) {
        set_error("dgtta_mind_ssc_fwd: noise_mode %d not supported", noise_mode);
        return DGTTA_EUNSUPPORTED;
    }
    if (noise_mode != DGTTA_NOISE_NONE && !noise_dev) { set_error("dgtta_mind_ssc_fwd: noise tensor / scratch missing"); return DGTTA_ENULL; }
    if (noise_mode == DGTTA_NOISE_PHILOX) {
        // regenerate torch.randn_like(edge_selection) for generator state (seed, offset) into the caller's scratch,
        // then run the streamed-noise kernel on it
        int dev = 0, mt = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&mt, cudaDevAttrMaxThreadsPerMultiProcessor, dev) != cudaSuccess) {
            set_error("dgtta_mind_ssc_fwd: device query failed");
            return (int)cudaGetLastError();
        }
        const int rc = philox_normal_fill(const_cast<float *>(noise_dev), (unsigned long long)B * 12 * D * H * W, philox_seed,
                                          philox_offset, sm_count(), mt, (cudaStream_t)stream);
        if (rc) return rc;
        noise_mode = DGTTA_NOISE_TENSOR;
    }
    if ((uintptr_t)workspace_dev & 15) { set_error("dgtta_mind_ssc_fwd: workspace must be 16-byte aligned"); return DGTTA_EWORKSPACE; }
    MindArgs a;
    a.img = img_dev; a.out = out_dev; a.noise = noise_dev; a.in_scale = in_scale_dev;
    a.workspace = workspace_dev; a.workspace_bytes = workspace_bytes;
    a.B = B; a.D = D; a.H = H; a.W = W; a.delta = delta; a.ntaps = ntaps;
    a.noise_mode = noise_mode; a.rw = randn_weighting;
    for (int i = 0; i < 9; ++i) a.taps[i] = i < ntaps ? taps_host[i] : 0.f;
    if (mind_fast_supported(a)) return mind_fast_launch(a, (cudaStream_t)stream);
    return mind_general_launch(a, (cudaStream_t)stream);
}
