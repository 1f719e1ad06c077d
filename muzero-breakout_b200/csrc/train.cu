// The two non-convolution ends of the training step (first slice):
//   mz_loss  — loss_fn (train_torch.py:33-66) with ScalarTransforms.supports_representation (utils.py:30-64):
//              the three batch-mean KL divergences, the total loss AND its gradient w.r.t. the three logit tensors
//              (what loss.backward() hands to the networks, train_torch.py:515), one launch.
//   mz_adam  — torch.optim.Adam(lr, weight_decay) of networks.py:268 over one flat parameter buffer:
//              HBM-bound, 28 B per parameter (read p, g, m, v; write p, m, v), 128-bit accesses.
// The convolution backward passes of the three networks are not built.
#include "common.cuh"

namespace mzb {

constexpr int kLossThreads = 128;
constexpr int kMaxSupports = 32;

struct loss_args {
    int rows, n_sup, n_act;
    float inv_rows, grad_scale;  // 1 / rows ; (1 / K) / rows
    const float *supports;
    const float *pred_reward, *pred_value, *pred_policy;
    const float *obs_reward, *value_target, *visits;
    float *d_reward, *d_value, *d_policy;
    double *partials;        // [gridDim.x][3]
    unsigned int *ticket;    // [1], zero on entry, reset by the last CTA
    float *losses;           // [4] total, reward, value, policy
    float inv_K;
};

// ScalarTransforms._invertible_transform_normal_to_compact (utils.py:21-24), every op separately rounded
__device__ __forceinline__ float to_compact(float x)
{
    const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    const float r = __fadd_rn(__fsub_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.f)), 1.f), __fmul_rn(0.001f, x));
    return __fmul_rn(sgn, r);
}

// one support-distribution head: KL(target || softmax(z)) summed over the row + gradient of it w.r.t. z (scaled)
// target = two-hot supports_representation of `scalar` (utils.py:44-62)
__device__ __forceinline__ float support_head(const loss_args &a, const float *__restrict__ z_row, float scalar, float *__restrict__ d_row)
{
    const int n = a.n_sup;
    float z[kMaxSupports];
    float zmax = -INFINITY;
    for (int i = 0; i < n; ++i) {
        z[i] = z_row[i];
        zmax = fmaxf(zmax, z[i]);
    }
    float se = 0.f;
    for (int i = 0; i < n; ++i) se += expf(z[i] - zmax);
    const float lse = logf(se);

    const float t = to_compact(scalar);
    int cnt = 0;                                   // searchsorted(supports, t, right=True): supports <= t
    for (int i = 0; i < n; ++i) cnt += (a.supports[i] <= t) ? 1 : 0;
    int lo = cnt - 1;
    lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
    const float s_lo = a.supports[lo], s_hi = a.supports[lo + 1];
    const float p_lo = __fdiv_rn(__fsub_rn(s_hi, t), __fadd_rn(__fsub_rn(s_hi, s_lo), 1e-10f));
    const float p_hi = __fsub_rn(1.f, p_lo);
    const float tsum = p_lo + p_hi;

    // xlogy(t, t) - t * logp over the two non-zero targets (a zero target contributes exactly 0)
    const float lp_lo = (z[lo] - zmax) - lse, lp_hi = (z[lo + 1] - zmax) - lse;
    float kl = 0.f;
    kl += (p_lo == 0.f) ? 0.f : (p_lo * logf(p_lo) - p_lo * lp_lo);
    kl += (p_hi == 0.f) ? 0.f : (p_hi * logf(p_hi) - p_hi * lp_hi);

    if (d_row != nullptr) {
        const float inv_se = 1.f / se;
        for (int i = 0; i < n; ++i) {
            const float sm = expf(z[i] - zmax) * inv_se;
            const float ti = (i == lo) ? p_lo : ((i == lo + 1) ? p_hi : 0.f);
            d_row[i] = a.grad_scale * (sm * tsum - ti);
        }
    }
    return kl;
}

// policy head: target = visit_counts / sum (train_torch.py:58), a row whose visits sum to 0 gives NaN like the reference
__device__ __forceinline__ float policy_head(const loss_args &a, const float *__restrict__ z_row, const float *__restrict__ v_row, float *__restrict__ d_row)
{
    const int n = a.n_act;
    float z[kMaxSupports], t[kMaxSupports];
    float zmax = -INFINITY, vsum = 0.f;
    for (int i = 0; i < n; ++i) {
        z[i] = z_row[i];
        zmax = fmaxf(zmax, z[i]);
        vsum += v_row[i];
    }
    float se = 0.f, tsum = 0.f;
    for (int i = 0; i < n; ++i) {
        se += expf(z[i] - zmax);
        t[i] = __fdiv_rn(v_row[i], vsum);
        tsum += t[i];
    }
    const float lse = logf(se), inv_se = 1.f / se;
    float kl = 0.f;
    for (int i = 0; i < n; ++i) {
        const float lp = (z[i] - zmax) - lse;
        kl += (t[i] == 0.f) ? 0.f : (t[i] * logf(t[i]) - t[i] * lp);
        if (d_row != nullptr) d_row[i] = a.grad_scale * (expf(z[i] - zmax) * inv_se * tsum - t[i]);
    }
    return kl;
}

__global__ void __launch_bounds__(kLossThreads) loss_kernel(const loss_args a)
{
    __shared__ double red[3][kLossThreads / 32];
    __shared__ bool last;
    const int row = blockIdx.x * kLossThreads + threadIdx.x;
    double part[3] = {0.0, 0.0, 0.0};
    if (row < a.rows) {
        part[0] = support_head(a, a.pred_reward + (size_t)row * a.n_sup, a.obs_reward[row], a.d_reward ? a.d_reward + (size_t)row * a.n_sup : nullptr);
        part[1] = support_head(a, a.pred_value + (size_t)row * a.n_sup, a.value_target[row], a.d_value ? a.d_value + (size_t)row * a.n_sup : nullptr);
        part[2] = policy_head(a, a.pred_policy + (size_t)row * a.n_act, a.visits + (size_t)row * a.n_act, a.d_policy ? a.d_policy + (size_t)row * a.n_act : nullptr);
    }
    // fixed-order reduction: lanes (shuffle tree), warps, then CTAs in index order by the last CTA to arrive -> deterministic
    for (int h = 0; h < 3; ++h) {
        double v = part[h];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[h][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int h = 0; h < 3; ++h) {
            double v = 0.0;
            for (int w = 0; w < kLossThreads / 32; ++w) v += red[h][w];
            a.partials[(size_t)blockIdx.x * 3 + h] = v;
        }
        __threadfence();
        last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    double tot[3] = {0.0, 0.0, 0.0};
    for (unsigned b = 0; b < gridDim.x; ++b)
        for (int h = 0; h < 3; ++h) tot[h] += ((volatile double *)a.partials)[(size_t)b * 3 + h];
    const float rl = (float)(tot[0] * (double)a.inv_rows), vl = (float)(tot[1] * (double)a.inv_rows), pl = (float)(tot[2] * (double)a.inv_rows);
    a.losses[0] = __fmul_rn(a.inv_K, __fadd_rn(__fadd_rn(rl, vl), pl));   // (1/K) * (reward_loss + value_loss + policy_loss), :66
    a.losses[1] = rl;
    a.losses[2] = vl;
    a.losses[3] = pl;
    *a.ticket = 0;   // ready for the next launch / graph replay
}

struct adam_consts {
    float wd, one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step_size;
};

// torch/optim/adam.py _single_tensor_adam, op by op: grad.add(param, alpha=wd); exp_avg.lerp_(grad, 1-b1);
// exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1-b2); denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps);
// param.addcdiv_(exp_avg, denom, value=-step_size).  Torch's CPU kernels contract add(alpha) / lerp / addcmul into FMAs and evaluate
// addcdiv as (value * exp_avg) / denom (probed against torch 2.11, tests/golden/gen_golden.py gen_train): with these the moments are
// bit-identical to torch's and the parameters within 1 ulp.
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const adam_consts &c)
{
    if (c.wd != 0.f) g = __fmaf_rn(c.wd, p, g);
    m = __fmaf_rn(c.one_minus_b1, __fsub_rn(g, m), m);
    v = __fmaf_rn(__fmul_rn(c.one_minus_b2, g), g, __fmul_rn(v, c.b2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(c.neg_step_size, m), denom));
}

__global__ void __launch_bounds__(256) adam_kernel(long long n4, long long n, float *__restrict__ param, const float *__restrict__ grad,
                                                   float *__restrict__ exp_avg, float *__restrict__ exp_avg_sq, const adam_consts c)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4 *>(param)[i];
        const float4 g = __ldcs(reinterpret_cast<const float4 *>(grad) + i);   // gradients are read once
        float4 m = reinterpret_cast<float4 *>(exp_avg)[i];
        float4 v = reinterpret_cast<float4 *>(exp_avg_sq)[i];
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        reinterpret_cast<float4 *>(param)[i] = p;
        reinterpret_cast<float4 *>(exp_avg)[i] = m;
        reinterpret_cast<float4 *>(exp_avg_sq)[i] = v;
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) adam_one(param[i], grad[i], exp_avg[i], exp_avg_sq[i], c);
}

// mz_adam_dev: the same update with the step count kept on the DEVICE, so that a CUDA-graph replay of a whole training step advances it.
// state: int32 step count (updates done so far) at byte 0, this update's constants at byte 16.
struct adam_hyper {
    double lr, beta1, beta2, eps, wd;
};
__global__ void adam_prep_kernel(void *state, const adam_hyper h)
{
    int *step = reinterpret_cast<int *>(state);
    adam_consts *c = reinterpret_cast<adam_consts *>(reinterpret_cast<char *>(state) + 16);
    const int t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow(h.beta1, (double)t), bc2 = 1.0 - pow(h.beta2, (double)t);
    c->wd = (float)h.wd;
    c->one_minus_b1 = (float)(1.0 - h.beta1);
    c->b2 = (float)h.beta2;
    c->one_minus_b2 = (float)(1.0 - h.beta2);
    c->bc2_sqrt = (float)sqrt(bc2);
    c->eps = (float)h.eps;
    c->neg_step_size = (float)(-(h.lr / bc1));
}
__global__ void __launch_bounds__(256) adam_dev_kernel(long long n4, long long n, float *__restrict__ param, const float *__restrict__ grad,
                                                       float *__restrict__ exp_avg, float *__restrict__ exp_avg_sq, const void *state)
{
    const adam_consts c = *reinterpret_cast<const adam_consts *>(reinterpret_cast<const char *>(state) + 16);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4 *>(param)[i];
        const float4 g = __ldcs(reinterpret_cast<const float4 *>(grad) + i);
        float4 m = reinterpret_cast<float4 *>(exp_avg)[i];
        float4 v = reinterpret_cast<float4 *>(exp_avg_sq)[i];
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        reinterpret_cast<float4 *>(param)[i] = p;
        reinterpret_cast<float4 *>(exp_avg)[i] = m;
        reinterpret_cast<float4 *>(exp_avg_sq)[i] = v;
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) adam_one(param[i], grad[i], exp_avg[i], exp_avg_sq[i], c);
}

}  // namespace mzb

extern "C" {

size_t mz_loss_scratch_bytes(int rows)
{
    if (rows <= 0) return 0;
    const size_t blocks = ((size_t)rows + mzb::kLossThreads - 1) / mzb::kLossThreads;
    return blocks * 3 * sizeof(double) + 16;
}

int mz_loss(int rows, int K, int n_supports, int n_actions, const float *supports, const float *pred_reward, const float *pred_value,
            const float *pred_policy, const float *observed_reward, const float *value_target, const float *visit_counts,
            float *losses, float *d_reward, float *d_value, float *d_policy, void *scratch, void *stream)
{
    using namespace mzb;
    MZB_CHECK_ARG(rows > 0 && K > 0, "rows and K must be positive");
    MZB_CHECK_ARG(n_supports >= 2 && n_supports <= kMaxSupports, "n_supports must be in [2, 32]");
    MZB_CHECK_ARG(n_actions >= 1 && n_actions <= kMaxSupports, "n_actions must be in [1, 32]");
    MZB_CHECK_ARG(supports && pred_reward && pred_value && pred_policy && observed_reward && value_target && visit_counts && losses && scratch,
                  "null pointer");
    MZB_CHECK_ARG(((uintptr_t)scratch & 7) == 0, "scratch must be 8-byte aligned");
    const int blocks = (rows + kLossThreads - 1) / kLossThreads;
    loss_args a;
    a.rows = rows; a.n_sup = n_supports; a.n_act = n_actions;
    a.inv_rows = 1.f / (float)rows;
    a.inv_K = (float)(1.0 / (double)K);
    a.grad_scale = (float)((1.0 / (double)K) / (double)rows);
    a.supports = supports;
    a.pred_reward = pred_reward; a.pred_value = pred_value; a.pred_policy = pred_policy;
    a.obs_reward = observed_reward; a.value_target = value_target; a.visits = visit_counts;
    a.d_reward = d_reward; a.d_value = d_value; a.d_policy = d_policy;
    a.partials = (double *)scratch;
    a.ticket = (unsigned int *)((char *)scratch + (size_t)blocks * 3 * sizeof(double));   // zeroed by the caller once, kept zero by the kernel
    a.losses = losses;
    loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(a);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_adam(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, double lr, double beta1, double beta2,
            double eps, double weight_decay, int step, void *stream)
{
    using namespace mzb;
    MZB_CHECK_ARG(n > 0 && step >= 1, "n must be positive and step >= 1 (the count AFTER this update, like torch's state['step'])");
    MZB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "null pointer");
    MZB_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "buffers must be 16-byte aligned");
    // scalars in double like torch/optim/adam.py, rounded to fp32 where they enter the tensor ops
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    adam_consts c;
    c.wd = (float)weight_decay;
    c.one_minus_b1 = (float)(1.0 - beta1);
    c.b2 = (float)beta2;
    c.one_minus_b2 = (float)(1.0 - beta2);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = (float)eps;
    c.neg_step_size = (float)(-(lr / bc1));
    const long long n4 = n / 4;
    long long want = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;       // 8 CTAs of 256 threads per SM, grid-stride
    const int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n4, n, param, grad, exp_avg, exp_avg_sq, c);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_adam_dev(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, double lr, double beta1, double beta2,
                double eps, double weight_decay, void *state, void *stream)
{
    using namespace mzb;
    MZB_CHECK_ARG(n > 0 && param && grad && exp_avg && exp_avg_sq && state, "bad argument");
    MZB_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)state) & 15) == 0, "buffers must be 16-byte aligned");
    adam_hyper h{lr, beta1, beta2, eps, weight_decay};
    adam_prep_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, h);
    MZB_LAUNCH_CHECK();
    const long long n4 = n / 4;
    long long want = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    const int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    adam_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n4, n, param, grad, exp_avg, exp_avg_sq, state);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
