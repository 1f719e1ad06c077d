// The glue of one acting move on the device (the "next" rows of SURVEY.md section 8f): observation-history
// ring -> representation-network input, and temperature sampling of the action from the root visit counts.
//
// Replaces (behaviour, not code) the per-environment Python loops of the reference's train_torch.py:
//   _prepare_mcts_input :259-277 + _encode_actions :279-293 + _pad_initial_state :313-332  (rep-net input)
//   temperature sampling :192-198 (visit_counts ** (1/T), normalise, Categorical.sample per env)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T cvt(float v);
template <> __device__ __forceinline__ float cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt<__half>(float v) { return __float2half_rn(v); }

// One thread per (env, pixel): writes the 64 channels of that pixel contiguously (channels-last).
//   channels 0..30  the last 31 frames appended to the trajectory, oldest first (ObservationTrajectory.get_states()[-31:])
//   channel  31     the current frame (train_torch.py:272 concatenates it again -- it is also the newest appended one)
//   channels 32..63 the last 32 actions / 3 as constant planes, oldest first (:268-269, :291-292)
// frames: float32 [slots][B][320] ring, `head` = slot of the newest appended frame; acts: int32 [slots][B] ring.
template <typename T>
__global__ void __launch_bounds__(256)
rep_input_kernel(int B, int slots, const float *__restrict__ frames, int head, const float *__restrict__ cur,
                 const int *__restrict__ acts, int ahead, T *__restrict__ out)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * 320) return;
    const int b = (int)(t / 320), p = (int)(t - (size_t)b * 320);
    T *o = out + t * 64;
#pragma unroll 1
    for (int k = 0; k < 31; ++k) {
        const int slot = (head - 30 + k + 2 * slots) % slots;
        o[k] = cvt<T>(frames[((size_t)slot * B + b) * 320 + p]);
    }
    o[31] = cvt<T>(cur[(size_t)b * 320 + p]);
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        const int slot = (ahead - 31 + k + 2 * slots) % slots;
        o[32 + k] = cvt<T>(__fdiv_rn((float)acts[(size_t)slot * B + b], 3.0f));
    }
}

// visit_counts ** (1/T) -> probabilities -> one categorical draw per env from the counter-based stream
// u32(seed, env, step) (the reference uses torch's global generator, train_torch.py:192-198).
__global__ void sample_actions_kernel(int B, const long long *__restrict__ visits, float inv_temperature, unsigned long long seed,
                                      unsigned step, long long *__restrict__ action, int *__restrict__ act_slot, float *__restrict__ probs_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float w[3], sum = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float n = (float)visits[b * 3 + a];
        w[a] = n > 0.0f ? powf(n, inv_temperature) : 0.0f;          // 0 ** x = 0 for x > 0
        sum += w[a];
    }
    const float u = (float)(mzb::mz_rng_u32(seed, (uint32_t)b, step) >> 8) * (1.0f / 16777216.0f);   // [0,1)
    const float p0 = w[0] / sum, p1 = w[1] / sum, p2 = w[2] / sum;
    int a = u < p0 ? 0 : (u < p0 + p1 ? 1 : 2);
    if (w[a] == 0.0f) a = w[2] > 0.0f ? 2 : (w[1] > 0.0f ? 1 : 0);      // never pick an unvisited action
    action[b] = a;
    if (act_slot) act_slot[b] = a;
    if (probs_out) { probs_out[b * 3] = p0; probs_out[b * 3 + 1] = p1; probs_out[b * 3 + 2] = p2; }
}

}  // namespace

extern "C" {

int mz_rep_input(int B, int slots, const float *frames, int head, const float *cur, const int32_t *acts, int ahead, void *out,
                 int dtype, void *stream)
{
    MZB_CHECK_ARG(B > 0 && slots >= 32 && frames && cur && acts && out, "bad argument");
    MZB_CHECK_ARG(head >= 0 && head < slots && ahead >= 0 && ahead < slots, "ring index out of range");
    const size_t total = (size_t)B * 320;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == MZ_F32) rep_input_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(B, slots, frames, head, cur, acts, ahead, (float *)out);
    else if (dtype == MZ_BF16) rep_input_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(B, slots, frames, head, cur, acts, ahead, (__nv_bfloat16 *)out);
    else if (dtype == MZ_F16) rep_input_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(B, slots, frames, head, cur, acts, ahead, (__half *)out);
    else { mzb::set_error("mz_rep_input: unknown dtype %d", dtype); return -1; }
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_sample_actions(int B, const int64_t *visits, double temperature, uint64_t seed, uint32_t step, int64_t *action,
                      int32_t *act_slot, float *probs_out, void *stream)
{
    MZB_CHECK_ARG(B > 0 && visits && action && temperature > 0.0, "bad argument");
    sample_actions_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, (const long long *)visits, (float)(1.0 / temperature), seed, step,
                                                                             (long long *)action, act_slot, probs_out);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
