// N(0,1) field of MIND3D's edge noise (dg_tta/mind.py:150, torch.randn_like(edge_selection)) regenerated on the
// device, bit-identical to what torch's CUDA generator writes for the same (seed, offset).
//
// torch (ATen/native/cuda/DistributionTemplates.h: calc_execution_policy + distribution_elementwise_grid_stride_kernel
// with curand_normal4) launches G = min(SMs * maxThreadsPerSM / 256, ceil(numel / 256)) blocks of 256 threads; with
// T = 256 G, thread idx < T initialises Philox4x32-10 with (key = seed, subsequence = idx, offset) and its j-th
// engine call fills elements  idx + T (4 j + ii), ii = 0..3  with the four Box-Muller normals of that call.  For an
// offset that is a multiple of 4 (torch only ever advances it by multiples of 4) the j-th call is the Philox block
// with counter (offset / 4 + j, idx).  Here every (idx, j) pair is an independent work item: no per-thread generator
// state, no wasted look-ahead block, round keys in the constant bank, 32-bit offsets, bounds checks only in the last
// engine call, four coalesced 128-byte stores per warp and pair.  The Box-Muller transform is the CUDA toolkit's own
// inline device function (curand_normal.h: the float arithmetic torch runs); tests/test_philox_gpu.py checks the
// stream bitwise against torch.randn.
#include <curand_kernel.h>

#include "common.cuh"

namespace dgtta {

namespace philox {

constexpr int BLOCK = 256;   // torch's block_size_bound: part of the stream's definition, not a tuning knob
constexpr int JB = 4;        // engine calls per thread

struct Keys {
    unsigned k[10][2];   // Philox4x32-10 round keys (key + r * Weyl constants): uniform, so they live in the constant bank
};

// One Philox4x32-10 block (the integer function behind curand_Philox4x32_10; Salmon et al., SC'11): per round two
// 32x32->64 multiplies and two 3-input xors.
__device__ __forceinline__ uint4 philox10(uint4 c, const Keys &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c.x;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c.z;
        c = make_uint4((unsigned)(p1 >> 32) ^ c.y ^ K.k[r][0], (unsigned)p1, (unsigned)(p0 >> 32) ^ c.w ^ K.k[r][1], (unsigned)p0);
    }
    return c;
}

template <bool TAIL>
__device__ __forceinline__ void one_call(float *__restrict__ out, unsigned numel, unsigned T, unsigned idx, unsigned long long c,
                                         int j, const Keys &K)
{
    const uint4 r = philox10(make_uint4((unsigned)c, (unsigned)(c >> 32), idx, 0u), K);
    const float2 a = _curand_box_muller(r.x, r.y);   // curand_normal.h: the float arithmetic torch's curand_normal4 runs
    const float2 b = _curand_box_muller(r.z, r.w);
    const unsigned li = idx + T * 4u * (unsigned)j;   // numel < 2^31 and li < numel + 4T: 32-bit offsets suffice
    if (!TAIL || li < numel) __stcs(out + li, a.x);
    if (!TAIL || li + T < numel) __stcs(out + li + T, a.y);
    if (!TAIL || li + 2u * T < numel) __stcs(out + li + 2u * T, b.x);
    if (!TAIL || li + 3u * T < numel) __stcs(out + li + 3u * T, b.y);
}

// state_dev == NULL: round keys and first counter are launch parameters.  state_dev != NULL (CUDA-graph mode): the
// generator state {seed, offset} is read from device memory at run time — the way torch itself feeds Philox under
// capture — so one captured launch draws a fresh field on every replay (the host updates the two words, e.g. through a
// captured memcpy from pinned memory, and advances torch's generator by dgtta_philox_normal_offset_increment).
template <bool DEV_STATE>
__global__ void __launch_bounds__(BLOCK) normal_fill_kernel(float *__restrict__ out, unsigned numel, unsigned T,
                                                            const __grid_constant__ Keys K0, unsigned long long ctr0_, int J,
                                                            const unsigned long long *__restrict__ state_dev)
{
    Keys K = K0;   // DEV_STATE == false: never written, stays in the constant bank
    unsigned long long ctr0 = ctr0_;
    if (DEV_STATE) {
        const unsigned long long seed = __ldg(state_dev), offset = __ldg(state_dev + 1);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            K.k[r][0] = (unsigned)seed + (unsigned)r * 0x9E3779B9u;
            K.k[r][1] = (unsigned)(seed >> 32) + (unsigned)r * 0xBB67AE85u;
        }
        ctr0 = offset >> 2;
    }
    const unsigned idx = blockIdx.x * BLOCK + threadIdx.x;   // < T
    const int j0 = blockIdx.y * JB;
    if (j0 + JB < J) {
        // every engine call but the last writes four in-range elements: no bounds checks
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) one_call<false>(out, numel, T, idx, ctr0 + (unsigned long long)(j0 + jj), j0 + jj, K);
    } else {
        for (int j = j0; j < J; ++j) one_call<true>(out, numel, T, idx, ctr0 + (unsigned long long)j, j, K);
    }
}

}  // namespace philox

void preload_philox() { DGTTA_TOUCH(philox::normal_fill_kernel<false>); DGTTA_TOUCH(philox::normal_fill_kernel<true>); }

// shared with mind_ssc.cu (DGTTA_NOISE_PHILOX)
int philox_normal_fill(float *out, unsigned long long numel, unsigned long long seed, unsigned long long offset, int sms,
                       int max_threads_per_sm, cudaStream_t stream, const unsigned long long *state_dev = nullptr)
{
    if (numel == 0) return 0;
    if (!out) { set_error("dgtta_philox_normal_fill: null pointer"); return DGTTA_ENULL; }
    if (offset & 3ull) { set_error("dgtta_philox_normal_fill: generator offset %llu is not a multiple of 4", offset); return DGTTA_EUNSUPPORTED; }
    if (numel >= (1ull << 31)) {
        // torch splits tensors that need 64-bit indexing into several launches with their own offsets
        set_error("dgtta_philox_normal_fill: %llu elements need torch's split launches (not reproduced)", numel);
        return DGTTA_EUNSUPPORTED;
    }
    if (sms <= 0 || max_threads_per_sm < philox::BLOCK) { set_error("dgtta_philox_normal_fill: bad device properties"); return DGTTA_EINVAL; }
    unsigned long long G = (numel + philox::BLOCK - 1) / philox::BLOCK;
    const unsigned long long cap = (unsigned long long)sms * (unsigned long long)(max_threads_per_sm / philox::BLOCK);
    if (G > cap) G = cap;
    const unsigned long long T = G * philox::BLOCK;
    const int J = (int)((numel - 1) / (T * 4ull) + 1);
    const dim3 grid((unsigned)G, (unsigned)((J + philox::JB - 1) / philox::JB));
    philox::Keys K;
    for (int r = 0; r < 10; ++r) {
        K.k[r][0] = (unsigned)seed + (unsigned)r * 0x9E3779B9u;
        K.k[r][1] = (unsigned)(seed >> 32) + (unsigned)r * 0xBB67AE85u;
    }
    if (state_dev) philox::normal_fill_kernel<true><<<grid, philox::BLOCK, 0, stream>>>(out, (unsigned)numel, (unsigned)T, K, 0ull, J, state_dev);
    else philox::normal_fill_kernel<false><<<grid, philox::BLOCK, 0, stream>>>(out, (unsigned)numel, (unsigned)T, K, offset / 4ull, J, nullptr);
    return check_launch("philox_normal_fill_kernel");
}

}  // namespace dgtta

extern "C" int dgtta_philox_normal_fill(float *out_dev, uint64_t numel, uint64_t philox_seed, uint64_t philox_offset,
                                        int sm_count, int max_threads_per_sm, dgtta_stream_t stream)
{
    return dgtta::philox_normal_fill(out_dev, numel, philox_seed, philox_offset, sm_count, max_threads_per_sm,
                                     (cudaStream_t)stream);
}

extern "C" int dgtta_philox_normal_fill_graphsafe(float *out_dev, uint64_t numel, const uint64_t *seed_offset_dev, int sm_count,
                                                  int max_threads_per_sm, dgtta_stream_t stream)
{
    if (!seed_offset_dev) { dgtta::set_error("dgtta_philox_normal_fill_graphsafe: null state"); return DGTTA_ENULL; }
    return dgtta::philox_normal_fill(out_dev, numel, 0, 0, sm_count, max_threads_per_sm, (cudaStream_t)stream,
                                     reinterpret_cast<const unsigned long long *>(seed_offset_dev));
}

extern "C" uint64_t dgtta_philox_normal_offset_increment(uint64_t numel, int sm_count, int max_threads_per_sm)
{
    if (numel == 0 || sm_count <= 0 || max_threads_per_sm < dgtta::philox::BLOCK) return 0;
    uint64_t G = (numel + dgtta::philox::BLOCK - 1) / dgtta::philox::BLOCK;
    const uint64_t cap = (uint64_t)sm_count * (uint64_t)(max_threads_per_sm / dgtta::philox::BLOCK);
    if (G > cap) G = cap;
    return ((numel - 1) / (dgtta::philox::BLOCK * G * 4) + 1) * 4;   // calc_execution_policy: engine calls * 4
}
