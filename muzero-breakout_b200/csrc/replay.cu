// Device-resident replay buffer: trajectory store, n-step value targets, FIFO sample ring
// and minibatch gather.
//
// Replaces (behaviour, not code) the reference's replay_buffer.py:
//   ObservationTrajectory.add_observation :17-35 + train_torch.py:204-208,313-332  (per-env Python lists, 32 padded rows)
//   ReplayBuffer.save_observation_trajectory :96-165   (per-sample re-tensorisation of the whole trajectory, nested
//                                                       Python loops for the value targets, list.pop(0) FIFO)
//   ReplayBuffer.get_batched_* :167-210, get_reward_sums :212-216
//
// Layout (all caller-owned HBM, described by rb_ring in include/mzb200.h).  The reference materialises every sample's
// 32-frame window (40 KB per sample, 2.4 GB at 60 000 samples); here a trajectory of T moves is stored ONCE as T+1
// consecutive entries of an entry ring (entry 0 = the initial frame, entry m+1 = move m: frame after the move, action,
// reward, visit counts, value) and a sample is 8 bytes of metadata (ring index of entry 0, start s, T) + its K value
// targets + the trajectory's reward sum.  Index algebra of the reference's padded lists (list index i, hist = 32):
//   actions[i] = entry[0].action (the padding action, 0 in acting) for i < 32, else entry[i-31]; same for rewards /
//   values / visits (their padding never reaches an output)               states[i] = entry[max(0, i-30)].frame
//   sample s: past actions i = s..s+31, frames i = s..s+31, future k: entry[s+1+k]
// Ring state (entries ever written, samples ever appended) lives on the device, so appending a whole batch of
// trajectories recorded by the on-device acting loop needs no host synchronisation.
//
// Kernels: rb_len (per-env length + sequential fp32 reward sum) -> rb_scan (one CTA: exclusive scans that place every
// trajectory in both rings, advances the ring state) -> rb_store (frames + scalars into the entry ring, 16-byte
// coalesced copies) -> rb_sample (metadata + value targets, one thread per (start, env), env fastest so the (T,B)
// records are read coalesced); rb_gather (one CTA per requested sample, 16-byte frame copies).  All HBM-bound copies.
#include "common.cuh"

namespace {

constexpr int FRAME = 320;            // 16 x 20 gray frame
constexpr int FRAME_V4 = FRAME / 4;   // 80 float4
constexpr int TD_STEPS = 10;          // replay_buffer.py:137

struct __align__(8) Plan {            // per-env placement, rb_plan_bytes(B) = B * sizeof(Plan)
    unsigned long long ent_base;      // global entry counter of entry 0
    unsigned long long smp_base;      // global sample counter of sample 0
    int len;                          // recorded moves
    int nsamples;                     // 0: trajectory not stored
    float rsum;
    int pad;
};

__global__ void __launch_bounds__(256)
rb_len_kernel(int B, int T, const float *__restrict__ reward, const uint8_t *__restrict__ recorded, const int32_t *__restrict__ lengths,
              int K, int min_length, int max_moves, Plan *__restrict__ plan, int32_t *__restrict__ status)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int len = 0;
    if (lengths) len = lengths[b];
    else for (int t = 0; t < T; ++t) len += recorded[(size_t)t * B + b] ? 1 : 0;     // recorded = "not done before the move": a prefix
    if (len < 0 || len > T || len > max_moves) { atomicOr(status, MZB_RB_ERR_BAD_LENGTH); len = 0; }
    float acc = 0.0f;                                                                  // replay_buffer.py:34, sequential fp32
    for (int t = 0; t < len; ++t) acc = __fadd_rn(acc, reward[(size_t)t * B + b]);
    Plan p;
    p.ent_base = p.smp_base = 0;
    p.len = len;
    p.nsamples = (len >= min_length && len >= K) ? len - K + 1 : 0;                    // range(length - K + 1), :106
    p.rsum = acc;
    p.pad = 0;
    plan[b] = p;
}

// One CTA: exclusive scan over the envs of (entries, samples) -> ring placement; advances the device ring state.
__global__ void __launch_bounds__(1024)
rb_scan_kernel(int B, Plan *__restrict__ plan, unsigned long long *__restrict__ state)
{
    __shared__ unsigned long long warp_e[32], warp_s[32];
    __shared__ unsigned long long carry_e, carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_e = state[0]; carry_s = state[1]; }
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + threadIdx.x;
        const int ns = b < B ? plan[b].nsamples : 0;
        const unsigned long long my_e = (b < B && ns > 0) ? (unsigned long long)plan[b].len + 1 : 0ull, my_s = (unsigned long long)ns;
        unsigned long long e = my_e, s = my_s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long pe = __shfl_up_sync(0xffffffffu, e, d), ps = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) { e += pe; s += ps; }
        }
        if (lane == 31) { warp_e[warp] = e; warp_s[warp] = s; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long we = warp_e[lane], ws = warp_s[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long pe = __shfl_up_sync(0xffffffffu, we, d), ps = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) { we += pe; ws += ps; }
            }
            warp_e[lane] = we; warp_s[lane] = ws;                                      // inclusive over warps
        }
        __syncthreads();
        const unsigned long long off_e = carry_e + (warp ? warp_e[warp - 1] : 0ull) + e - my_e;
        const unsigned long long off_s = carry_s + (warp ? warp_s[warp - 1] : 0ull) + s - my_s;
        if (b < B) { plan[b].ent_base = off_e; plan[b].smp_base = off_s; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_e += warp_e[31]; carry_s += warp_s[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { state[0] = carry_e; state[1] = carry_s; }
}

// grid (B, ceil((T+1)/4)), 320 threads = 4 entries x 80 float4 lanes.
__global__ void __launch_bounds__(320)
rb_store_kernel(rb_ring r, int B, int T, const long long *__restrict__ action, const float *__restrict__ reward, const float *__restrict__ value,
                const long long *__restrict__ visits, const float *__restrict__ frames, const float *__restrict__ init_frame,
                int pad_action, const Plan *__restrict__ plan)
{
    const int b = blockIdx.x;
    const int e = blockIdx.y * 4 + threadIdx.x / FRAME_V4, q = threadIdx.x % FRAME_V4;
    const Plan p = plan[b];
    if (p.nsamples <= 0 || e > p.len) return;
    // a batch larger than the ring: entries a later trajectory of this batch overwrites are never written (no race)
    if (p.ent_base + (unsigned long long)e + (unsigned long long)r.cap_entries < r.state[0]) return;
    const size_t slot = (size_t)((p.ent_base + (unsigned long long)e) % (unsigned long long)r.cap_entries);
    const float4 *src = reinterpret_cast<const float4 *>(e == 0 ? init_frame + (size_t)b * FRAME : frames + ((size_t)(e - 1) * B + b) * FRAME);
    reinterpret_cast<float4 *>(r.frame + slot * FRAME)[q] = __ldg(src + q);
    if (q == 0) {
        if (e == 0) {
            r.action[slot] = pad_action; r.reward[slot] = 0.0f; r.value[slot] = 0.0f;     // entry 0 carries the padding action
            r.visits[slot * 3] = r.visits[slot * 3 + 1] = r.visits[slot * 3 + 2] = 0.0f;
        } else {
            const size_t m = (size_t)(e - 1) * B + b;
            r.action[slot] = (int32_t)action[m]; r.reward[slot] = reward[m]; r.value[slot] = value[m];
            r.visits[slot * 3] = (float)visits[m * 3]; r.visits[slot * 3 + 1] = (float)visits[m * 3 + 1]; r.visits[slot * 3 + 2] = (float)visits[m * 3 + 2];
        }
    }
}

// One thread per (start s, env b), b fastest.  Value targets: replay_buffer.py:136-151, every op separately rounded.
__global__ void __launch_bounds__(256)
rb_sample_kernel(rb_ring r, int B, int nstarts, const float *__restrict__ reward, const float *__restrict__ value, const Plan *__restrict__ plan)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(t / B), b = (int)(t - (size_t)s * B);
    if (s >= nstarts) return;
    const Plan p = plan[b];
    if (s >= p.nsamples) return;
    if (p.smp_base + (unsigned long long)s + (unsigned long long)r.cap_samples < r.state[1]) return;   // evicted by this same batch (FIFO, :154-163)
    const size_t phys = (size_t)((p.smp_base + (unsigned long long)s) % (unsigned long long)r.cap_samples);
    const unsigned long long ent = p.ent_base % (unsigned long long)r.cap_entries;
    r.meta[phys] = ent | ((unsigned long long)s << 32) | ((unsigned long long)p.len << 48);
    r.reward_sum[phys] = p.rsum;
    const int K = r.K, len = p.len;
    for (int kk = 0; kk < K; ++kk) {
        const int m0 = s + kk, mb = m0 + TD_STEPS;          // current move, bootstrap move (list index - 32)
        float vt;
        int n;
        if (mb < len) { vt = __fmul_rn(value[(size_t)mb * B + b], r.gpow[K]); n = TD_STEPS; }
        else { vt = 0.0f; n = len - m0; }
        for (int k = 0; k < n; ++k) vt = __fadd_rn(vt, __fmul_rn(r.gpow[k], reward[(size_t)(m0 + k) * B + b]));
        r.target[phys * K + kk] = vt;
    }
}

// One CTA of 256 threads per requested sample.
__global__ void __launch_bounds__(256)
rb_gather_kernel(rb_ring r, int n, const long long *__restrict__ idx, long long *__restrict__ past_actions, long long *__restrict__ future_actions,
                 float *__restrict__ states, float *__restrict__ rewards, float *__restrict__ visit_counts, float *__restrict__ values,
                 float *__restrict__ value_buffer, float *__restrict__ reward_sums, int32_t *__restrict__ status)
{
    const int i = blockIdx.x;
    const unsigned long long total = r.state[1], cap = (unsigned long long)r.cap_samples;
    const long long length = (long long)(total < cap ? total : cap);
    long long li = idx[i];
    if (li < 0) li += length;                               // Python list indexing: -1 = newest
    const int K = r.K, hist = r.hist;
    const bool ok = li >= 0 && li < length;
    if (!ok && threadIdx.x == 0) atomicOr(status, MZB_RB_ERR_BAD_INDEX);          // the reference raises IndexError
    const size_t phys = (size_t)((total - (unsigned long long)length + (unsigned long long)(ok ? li : 0)) % cap);   // logical 0 = oldest live sample
    const unsigned long long meta = ok ? r.meta[phys] : 0ull;
    const unsigned long long ent = meta & 0xffffffffull, CE = (unsigned long long)r.cap_entries;
    const int s = (int)((meta >> 32) & 0xffff);
    auto slot_of = [&](int e) { return (size_t)((ent + (unsigned long long)e) % CE); };
    if (states) {
        float4 *dst = reinterpret_cast<float4 *>(states + (size_t)i * hist * FRAME);
        for (int w = threadIdx.x; w < hist * FRAME_V4; w += blockDim.x) {
            const int j = w / FRAME_V4, q = w - j * FRAME_V4;
            const int e = max(0, s + j - (hist - 2));                                   // states[i] = entry[max(0, i - 30)].frame
            dst[w] = ok ? reinterpret_cast<const float4 *>(r.frame + slot_of(e) * FRAME)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const int t = threadIdx.x;
    if (past_actions && t < hist) {
        const int li2 = s + t;                                                          // list index
        past_actions[(size_t)i * hist + t] = ok ? (long long)r.action[slot_of(li2 >= hist ? li2 - (hist - 1) : 0)] : 0ll;
    }
    if (t < K) {
        const size_t sl = slot_of(s + 1 + t);
        if (future_actions) future_actions[(size_t)i * K + t] = ok ? (long long)r.action[sl] : 0ll;
        if (rewards) rewards[(size_t)i * K + t] = ok ? r.reward[sl] : 0.0f;
        if (value_buffer) value_buffer[(size_t)i * K + t] = ok ? r.value[sl] : 0.0f;
        if (values) values[(size_t)i * K + t] = ok ? r.target[phys * K + t] : 0.0f;
        if (visit_counts) {
            for (int a = 0; a < 3; ++a) visit_counts[((size_t)i * K + t) * 3 + a] = ok ? r.visits[sl * 3 + a] : 0.0f;
        }
    }
    if (reward_sums && t == 0) reward_sums[i] = ok ? r.reward_sum[phys] : 0.0f;
}

// Representation-network input of a training minibatch: cat(states.view(n, hist, 16, 20), _encode_actions(past_actions)) along the plane axis
// (train_torch.py:392,500 with :279-293): planes 0..hist-1 = the frame window, planes hist..2*hist-1 = past_action / n_actions as constant planes.
__global__ void __launch_bounds__(256)
rb_gather_input_kernel(rb_ring r, int n, const long long *__restrict__ idx, float n_actions, float *__restrict__ out, int32_t *__restrict__ status)
{
    const int i = blockIdx.x;
    const unsigned long long total = r.state[1], cap = (unsigned long long)r.cap_samples;
    const long long length = (long long)(total < cap ? total : cap);
    long long li = idx[i];
    if (li < 0) li += length;
    const int hist = r.hist;
    const bool ok = li >= 0 && li < length;
    if (!ok && threadIdx.x == 0) atomicOr(status, MZB_RB_ERR_BAD_INDEX);
    const size_t phys = (size_t)((total - (unsigned long long)length + (unsigned long long)(ok ? li : 0)) % cap);
    const unsigned long long meta = ok ? r.meta[phys] : 0ull;
    const unsigned long long ent = meta & 0xffffffffull, CE = (unsigned long long)r.cap_entries;
    const int s = (int)((meta >> 32) & 0xffff);
    auto slot_of = [&](int e) { return (size_t)((ent + (unsigned long long)e) % CE); };
    float4 *dst = reinterpret_cast<float4 *>(out + (size_t)i * 2 * hist * FRAME);
    for (int w = threadIdx.x; w < 2 * hist * FRAME_V4; w += blockDim.x) {
        const int j = w / FRAME_V4, q = w - j * FRAME_V4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && j < hist) {
            v = reinterpret_cast<const float4 *>(r.frame + slot_of(max(0, s + j - (hist - 2))) * FRAME)[q];   // as rb_gather's states
        } else if (ok) {
            const int li2 = s + (j - hist);                                                                    // as rb_gather's past_actions
            const float a = __fdiv_rn((float)r.action[slot_of(li2 >= hist ? li2 - (hist - 1) : 0)], n_actions);
            v = make_float4(a, a, a, a);
        }
        __stcs(dst + w, v);                                                                                    // written once, read by the rep net
    }
}

int check_ring(const rb_ring *r)
{
    MZB_CHECK_ARG(r, "ring is NULL");
    MZB_CHECK_ARG(r->cap_samples > 0 && r->cap_entries > 0 && r->K >= 1 && r->K <= 15 && r->hist >= 2 && r->hist <= 256, "bad ring geometry");
    MZB_CHECK_ARG(r->frame && r->action && r->reward && r->value && r->visits && r->meta && r->reward_sum && r->target && r->state,
                  "ring array is NULL");
    return 0;
}

}  // namespace

extern "C" {

size_t rb_plan_bytes(int B) { return (size_t)(B > 0 ? B : 0) * sizeof(Plan); }

long long rb_entries_for(int cap_samples, int K, int max_moves)
{
    if (cap_samples <= 0 || K < 1 || max_moves < K) return -1;
    // live trajectories hold < cap + max_moves samples; a stored trajectory of T >= K moves has T+1 entries for T-K+1
    // samples, at most (K+1) entries per sample (T = K)
    return ((long long)cap_samples + max_moves) * (K + 1) + max_moves + 1;
}

int rb_append(const rb_ring *ring, int B, int T, const int64_t *action, const float *reward, const float *value, const int64_t *visits,
              const float *frames, const float *init_frame, int pad_action, const uint8_t *recorded, const int32_t *lengths, int min_length,
              int max_moves, void *plan, int32_t *status, void *stream)
{
    if (check_ring(ring)) return -1;
    MZB_CHECK_ARG(B > 0 && T >= 0 && plan && status, "bad argument");
    MZB_CHECK_ARG(T == 0 || (action && reward && value && visits && frames), "record array is NULL");
    MZB_CHECK_ARG(init_frame && (recorded || lengths), "init_frame / lengths missing");
    MZB_CHECK_ARG(max_moves > 0 && max_moves <= 65535 && T <= max_moves, "trajectory longer than max_moves");
    MZB_CHECK_ARG(ring->cap_entries >= rb_entries_for(ring->cap_samples, ring->K, max_moves),
                  "entry ring smaller than rb_entries_for(cap_samples, K, max_moves)");
    cudaStream_t st = (cudaStream_t)stream;
    Plan *pl = (Plan *)plan;
    rb_len_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, T, reward, recorded, lengths, ring->K, min_length, max_moves, pl, status);
    MZB_LAUNCH_CHECK();
    rb_scan_kernel<<<1, 1024, 0, st>>>(B, pl, (unsigned long long *)ring->state);
    MZB_LAUNCH_CHECK();
    if (T >= ring->K) {
        rb_store_kernel<<<dim3((unsigned)B, (unsigned)((T + 1 + 3) / 4)), 320, 0, st>>>(*ring, B, T, (const long long *)action, reward, value,
                                                                                        (const long long *)visits, frames, init_frame, pad_action, pl);
        MZB_LAUNCH_CHECK();
        const size_t total = (size_t)B * (size_t)(T - ring->K + 1);
        rb_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*ring, B, T - ring->K + 1, reward, value, pl);
        MZB_LAUNCH_CHECK();
    }
    return 0;
}

int rb_gather(const rb_ring *ring, int n, const int64_t *idx, int64_t *past_actions, int64_t *future_actions, float *states, float *rewards,
              float *visit_counts, float *values, float *value_buffer, float *reward_sums, int32_t *status, void *stream)
{
    if (check_ring(ring)) return -1;
    MZB_CHECK_ARG(n >= 0 && status && (n == 0 || idx), "bad argument");
    if (n == 0) return 0;
    rb_gather_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(*ring, n, (const long long *)idx, (long long *)past_actions,
                                                                     (long long *)future_actions, states, rewards, visit_counts, values,
                                                                     value_buffer, reward_sums, status);
    MZB_LAUNCH_CHECK();
    return 0;
}

int rb_gather_input(const rb_ring *ring, int n, const int64_t *idx, int n_actions, float *out, int32_t *status, void *stream)
{
    if (check_ring(ring)) return -1;
    MZB_CHECK_ARG(n >= 0 && n_actions > 0 && status && (n == 0 || (idx && out)), "bad argument");
    if (n == 0) return 0;
    rb_gather_input_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(*ring, n, (const long long *)idx, (float)n_actions, out, status);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
