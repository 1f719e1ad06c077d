// Vectorised Breakout environment for B200 (sm_100a): structure-of-arrays state in HBM, one warp per
// task of 2 environments (lane = env for the game logic, then the whole warp writes the 2 frames as one contiguous
// region), dense float32 frames written with coalesced 128-bit streaming stores.
//
// Replaces (behaviour, not code) environment/parallel_breakout.py of the reference:
//   reset :107-139, get_valid_actions :141-155, step :158-254, and the caller-side
//   convert_to_grayscale train_torch.py:334-358.  See include/mzb200.h for the C ABI and DESIGN.md
//   for the layout and the roofline (HBM-bound: 3840 B frame write per env-step dominates).
#include <string.h>

#include "common.cuh"

namespace {

constexpr int H = MZB_ENV_H, W = MZB_ENV_W;
constexpr int PADDLE_W = 6;
constexpr int FRAME_V4 = MZB_ENV_FRAME_FLOATS / 4;  // 240 float4 per frame
constexpr int GRAY_V4 = H * W / 4;                  // 80 float4 per gray frame
constexpr int WARPS_PER_BLOCK = 8;
constexpr int ENVS_PER_WARP = 2;       // envs per warp task (one 480-float4 period of the frame map).  Measured at 65 536 envs:
                                       // 32 -> 47.9 us, 8 -> 47.3, 4 -> 44.0, 2 -> 41.5 us (a plain 252 MB fill_ takes 37 us):
                                       // many small tasks hide the load latency of the logic phase and balance the SMs
constexpr int WORDS = 48;         // 3 planes x 16 row bitmasks per env
constexpr int WORDS_PAD = 49;     // +1: conflict-free when lane e writes words[e][r]

// ---- hdr packing (include/mzb200.h) ----
struct Hdr {
    int bx, by, px, vis, dx, dy;
    uint32_t rowmask;
};
__device__ __forceinline__ Hdr unpack(uint64_t h)
{
    Hdr s;
    uint32_t lo = (uint32_t)h;
    s.bx = lo & 31;
    s.by = (lo >> 5) & 15;
    s.px = (lo >> 9) & 15;
    s.vis = (lo >> 13) & 1;
    s.dx = (int)((lo >> 14) & 3) - 1;
    s.dy = (int)((lo >> 16) & 3) - 1;
    s.rowmask = (uint32_t)(h >> 32) & 0xFFFFu;
    return s;
}
__device__ __forceinline__ uint64_t pack(const Hdr &s)
{
    uint32_t lo = (uint32_t)s.bx | ((uint32_t)s.by << 5) | ((uint32_t)s.px << 9) | ((uint32_t)s.vis << 13) |
                  ((uint32_t)(s.dx + 1) << 14) | ((uint32_t)(s.dy + 1) << 16);
    return (uint64_t)lo | ((uint64_t)s.rowmask << 32);
}

__device__ __forceinline__ float4 nibble_to_float4(uint32_t nib)
{
    return make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f,
                       (nib & 8u) ? 1.0f : 0.0f);
}

// convert_to_grayscale, train_torch.py:350-356: clamp((p*0.3 + b*1.0) + k*0.6, 0, 1), each op rounded
__device__ __forceinline__ float gray_px(uint32_t p, uint32_t b, uint32_t k)
{
    float t = __fadd_rn(__fmul_rn(p ? 1.0f : 0.0f, 0.3f), __fmul_rn(b ? 1.0f : 0.0f, 1.0f));
    float v = __fadd_rn(t, __fmul_rn(k ? 1.0f : 0.0f, 0.6f));
    return fminf(fmaxf(v, 0.0f), 1.0f);
}

// Phase 2 of every frame-producing kernel: the warp's ENVS_PER_WARP frames are one contiguous region;
// every store instruction writes 512 contiguous bytes.  words = smem [32][WORDS_PAD] row bitmasks.
// 480 float4 (= 2 frames) is the period of the (lane -> plane,row,column-group) map, so the 15
// descriptors are computed once and reused for the ENVS_PER_WARP/2 frame pairs.
__device__ __forceinline__ void write_frames(const uint32_t *words, float *state_out, int env0, int B, int lane)
{
    float4 *out = reinterpret_cast<float4 *>(state_out) + (size_t)env0 * FRAME_V4;
    const int nenv = min(ENVS_PER_WARP, B - env0);
#pragma unroll
    for (int it = 0; it < 15; ++it) {
        const int j = it * 32 + lane;  // 0..479
        const int eo = j >= FRAME_V4 ? 1 : 0;
        const int i = j - eo * FRAME_V4;
        const int plane = i / 80;
        const int rem = i - plane * 80;
        const int rr = rem / 5;
        const int sh = (rem - rr * 5) * 4;
        const int widx = plane * 16 + rr;
#pragma unroll 4
        for (int pair = 0; pair < ENVS_PER_WARP / 2; ++pair) {
            const int e = pair * 2 + eo;
            if (e < nenv) {
                const uint32_t w = words[e * WORDS_PAD + widx];
                __stcs(out + pair * (2 * FRAME_V4) + j, nibble_to_float4((w >> sh) & 0xFu));
            }
        }
    }
}

__device__ __forceinline__ void write_gray(const uint32_t *words, float *gray_out, int env0, int B, int lane)
{
    float4 *out = reinterpret_cast<float4 *>(gray_out) + (size_t)env0 * GRAY_V4;
    const int nenv = min(ENVS_PER_WARP, B - env0);
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int j = it * 32 + lane;  // 0..159
        const int eo = j >= GRAY_V4 ? 1 : 0;
        const int i = j - eo * GRAY_V4;
        const int rr = i / 5;
        const int sh = (i - rr * 5) * 4;
#pragma unroll 4
        for (int pair = 0; pair < ENVS_PER_WARP / 2; ++pair) {
            const int e = pair * 2 + eo;
            if (e < nenv) {
                const uint32_t *w = words + e * WORDS_PAD;
                const uint32_t p = (w[rr] >> sh) & 0xFu, b = (w[16 + rr] >> sh) & 0xFu, k = (w[32 + rr] >> sh) & 0xFu;
                float4 v = make_float4(gray_px(p & 1, b & 1, k & 1), gray_px(p & 2, b & 2, k & 2),
                                       gray_px(p & 4, b & 4, k & 4), gray_px(p & 8, b & 8, k & 8));
                __stcs(out + pair * (2 * GRAY_V4) + j, v);
            }
        }
    }
}

// lane e publishes its env's 48 row words (paddle / ball planes derived from hdr; brick rows given)
__device__ __forceinline__ void publish_planes(uint32_t *w, const Hdr &s)
{
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        w[r] = (r == H - 1 && s.vis) ? (0x3Fu << s.px) : 0u;
        w[16 + r] = (r == s.by) ? (1u << s.bx) : 0u;
    }
}

template <bool kFrame, bool kGray>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
env_step_kernel(int B, uint64_t *__restrict__ hdr, uint32_t *__restrict__ bricks, const int64_t *__restrict__ action,
                uint8_t *__restrict__ done, float *__restrict__ next_state, float *__restrict__ reward,
                float *__restrict__ valid, float *__restrict__ gray, float4 rw, int32_t *__restrict__ status)
{
    __shared__ uint32_t s_words[WARPS_PER_BLOCK][ENVS_PER_WARP * WORDS_PAD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env0 = (blockIdx.x * WARPS_PER_BLOCK + warp) * ENVS_PER_WARP;
    if (env0 >= B) return;
    const int e = env0 + lane;
    uint32_t *w = s_words[warp] + (lane % ENVS_PER_WARP) * WORDS_PAD;

    if (lane < ENVS_PER_WARP && e < B) {
        Hdr s = unpack(hdr[e]);
        const int a = (int)action[e];
        const int din = done[e] ? 1 : 0;
        int err = (a < 0 || a > 2) ? MZB_ENV_ERR_BAD_ACTION : 0;

        // brick rows -> smem (coalesced across the warp: bricks[r][e])
#pragma unroll
        for (int r = 0; r < 16; ++r) w[32 + r] = ((s.rowmask >> r) & 1u) ? bricks[(size_t)r * B + e] : 0u;

        // paddle :177-186 (argmax of an all-zero row is 0)
        const int ppos = s.vis ? s.px : 0;
        int pnew = ppos + (a == 0 ? -1 : (a == 2 ? 1 : 0));
        pnew = max(0, min(W - PADDLE_W, pnew));
        // wall :195-196, move :198-199
        int dx = s.dx, dy = s.dy;
        if (s.bx + dx < 0 || s.bx + dx >= W) dx = -dx;
        int ny = s.by + dy;
        const int nx = s.bx + dx;
        // lost :202-209
        const int missed = ny >= H;
        float r = missed ? rw.z : 0.0f;
        int d = din | missed;
        uint32_t rowmask = s.rowmask;
        if (d) {
            rowmask = 0;
            dx = 0;
            dy = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) w[32 + q] = 0u;
        }
        if (missed) ny = 0;
        // ceiling :213-214
        if (ny < 0) { dy = -dy; ny = s.by; }
        // bricks :217-226
        const int old_dy = dy;
        const int cx = nx & ~1;
        const uint32_t row = w[32 + ny];
        const int hit = (row >> cx) & 1u;
        if (hit) dy = -old_dy;
        const uint32_t newrow = row & ~(3u << cx);
        if (newrow != row) {
            w[32 + ny] = newrow;
            bricks[(size_t)ny * B + e] = newrow;
            if (newrow == 0u) rowmask &= ~(1u << ny);
        }
        if (hit) { ny = s.by - old_dy; r = __fadd_rn(r, rw.y); }
        if (ny >= H) { err |= MZB_ENV_ERR_BALL_LEFT_GRID; ny = H - 1; }
        // paddle :229-239 (mask built from the new paddle position even for finished games)
        if (ny == H - 1 && nx >= pnew && nx < pnew + PADDLE_W) { dy = -dy; r = __fadd_rn(r, rw.x); }
        // terminal :246-250
        const int finished = rowmask == 0u;
        d |= finished;
        if (finished != missed) r = __fadd_rn(r, rw.w);

        s.bx = nx;
        s.by = ny & 15;  // row -1 wraps to row 15 (:243)
        s.px = pnew;
        s.vis = d ? 0 : 1;
        s.dx = dx;
        s.dy = dy;
        s.rowmask = d ? 0u : rowmask;
        hdr[e] = pack(s);
        done[e] = (uint8_t)d;
        reward[e] = r;
        valid[e * 3 + 0] = pnew == 0 ? 0.0f : 1.0f;      // get_valid_actions :141-155
        valid[e * 3 + 1] = 1.0f;
        valid[e * 3 + 2] = (pnew + PADDLE_W >= W) ? 0.0f : 1.0f;
        if (err) atomicOr(status, err);
        if (kFrame || kGray) publish_planes(w, s);
    }
    if (kFrame || kGray) __syncwarp();
    if (kFrame) write_frames(s_words[warp], next_state, env0, B, lane);
    if (kGray) write_gray(s_words[warp], gray, env0, B, lane);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
env_render_kernel(int B, const uint64_t *__restrict__ hdr, const uint32_t *__restrict__ bricks, float *__restrict__ state_out)
{
    __shared__ uint32_t s_words[WARPS_PER_BLOCK][ENVS_PER_WARP * WORDS_PAD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env0 = (blockIdx.x * WARPS_PER_BLOCK + warp) * ENVS_PER_WARP;
    if (env0 >= B) return;
    const int e = env0 + lane;
    uint32_t *w = s_words[warp] + (lane % ENVS_PER_WARP) * WORDS_PAD;
    if (lane < ENVS_PER_WARP && e < B) {
        Hdr s = unpack(hdr[e]);
#pragma unroll
        for (int r = 0; r < 16; ++r) w[32 + r] = ((s.rowmask >> r) & 1u) ? bricks[(size_t)r * B + e] : 0u;
        publish_planes(w, s);
    }
    __syncwarp();
    write_frames(s_words[warp], state_out, env0, B, lane);
}

__device__ __forceinline__ void init_env(int B, int e, uint64_t *hdr, uint32_t *bricks, int offset, int ball_x, int ball_h, int pick)
{
    Hdr s;
    s.px = W / 2 - PADDLE_W / 2 + offset;        // :120
    s.bx = ball_x;                               // :126
    s.by = ball_h < 0 ? H + ball_h : ball_h;     // :127-128 (negative row index)
    s.vis = 1;
    s.dx = pick ? 1 : -1;                        // :134-136
    s.dy = -1;                                   // :137
    s.rowmask = 0x7u;                            // brick rows 0..2 (:131; brick_rows is hard-coded 3, :79)
    hdr[e] = pack(s);
#pragma unroll
    for (int r = 0; r < 16; ++r) bricks[(size_t)r * B + e] = r < 3 ? 0xFFFFFu : 0u;
}

__global__ void env_reset_kernel(int B, uint64_t *hdr, uint32_t *bricks, const int64_t *offset, const int64_t *ball_x,
                                 const int64_t *ball_h, const int64_t *pick)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B) init_env(B, e, hdr, bricks, (int)offset[e], (int)ball_x[e], (int)ball_h[e], (int)pick[e]);
}

__global__ void env_reset_rng_kernel(int B, uint64_t *hdr, uint32_t *bricks, uint64_t seed, uint64_t episode)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    const uint64_t z = mzb::mix64(seed ^ mzb::mix64(episode * 0x9E3779B97F4A7C15ULL + (uint64_t)e + 1));
    // same ranges as torch.randint(-6,8) / (1,19) / (-3,-1) / (0,2), parallel_breakout.py:116,126,127,136
    init_env(B, e, hdr, bricks, -6 + (int)((z & 0xFFFF) % 14), 1 + (int)(((z >> 16) & 0xFFFF) % 18),
             -3 + (int)((z >> 32) & 1), (int)((z >> 33) & 1));
}

// dense -> SoA, one warp per env; lanes read the 240 float4 of the frame coalesced.
__global__ void __launch_bounds__(128)
env_ingest_kernel(int B, const float *__restrict__ state, const int64_t *__restrict__ ball_dx, const float *__restrict__ ball_dy,
                  uint64_t *__restrict__ hdr, uint32_t *__restrict__ bricks, int32_t *__restrict__ status)
{
    __shared__ uint32_t s_rows[4][WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * 4 + warp;
    if (e >= B) return;
    uint32_t *rows = s_rows[warp];
    for (int i = lane; i < WORDS; i += 32) rows[i] = 0u;
    __syncwarp();
    const float4 *src = reinterpret_cast<const float4 *>(state) + (size_t)e * FRAME_V4;
    int bad = 0;
    for (int i = lane; i < FRAME_V4; i += 32) {
        const float4 v = __ldg(src + i);
        const float f[4] = {v.x, v.y, v.z, v.w};
        uint32_t nib = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (f[q] == 1.0f) nib |= 1u << q;
            else if (f[q] != 0.0f) bad = 1;
        }
        const int plane = i / 80, rem = i - plane * 80, rr = rem / 5, sh = (rem - rr * 5) * 4;
        if (nib) atomicOr(&rows[plane * 16 + rr], nib << sh);
    }
    __syncwarp();
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        Hdr s = unpack(hdr[e]);
        // paddle: rows 0..14 empty, row 15 either empty or 6 consecutive cells
        const uint32_t prow = rows[15];
        for (int r = 0; r < 15; ++r) bad |= rows[r] != 0u;
        if (prow == 0u) { s.vis = 0; s.px = 0; }
        else {
            s.px = __ffs(prow) - 1;
            s.vis = 1;
            bad |= (prow != (0x3Fu << s.px)) || s.px > W - PADDLE_W;
        }
        int nball = 0;
        for (int r = 0; r < 16; ++r) {
            const uint32_t b = rows[16 + r];
            if (b) { nball += __popc(b); s.by = r; s.bx = __ffs(b) - 1; }
        }
        bad |= nball != 1;
        uint32_t mask = 0;
        for (int r = 0; r < 16; ++r) {
            bricks[(size_t)r * B + e] = rows[32 + r];
            if (rows[32 + r]) mask |= 1u << r;
        }
        s.rowmask = mask;
        if (ball_dx) { const int64_t d = ball_dx[e]; bad |= d < -1 || d > 1; s.dx = (int)max((int64_t)-1, min((int64_t)1, d)); }
        if (ball_dy) { const float d = ball_dy[e]; bad |= !(d == -1.0f || d == 0.0f || d == 1.0f); s.dy = d < 0.0f ? -1 : (d > 0.0f ? 1 : 0); }
        hdr[e] = pack(s);
        if (bad) atomicOr(status, MZB_ENV_ERR_BAD_STATE);
    }
}

__global__ void env_velocity_kernel(int B, const uint64_t *hdr, int64_t *ball_dx, float *ball_dy)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    const Hdr s = unpack(hdr[e]);
    ball_dx[e] = s.dx;
    ball_dy[e] = (float)s.dy;
}

__global__ void gray_kernel(size_t n4, const float4 *__restrict__ state, float4 *__restrict__ gray)
{
    // one float4 of output per thread: pixel group g of env b reads planes 0,1,2 at the same offset
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    const size_t b = t / GRAY_V4, g = t - b * GRAY_V4;
    const float4 *s = state + b * FRAME_V4 + g;
    const float4 p = __ldcs(s), q = __ldcs(s + GRAY_V4), k = __ldcs(s + 2 * GRAY_V4);
    auto px = [](float a, float c, float d) {
        float v = __fadd_rn(__fadd_rn(__fmul_rn(a, 0.3f), __fmul_rn(c, 1.0f)), __fmul_rn(d, 0.6f));
        return fminf(fmaxf(v, 0.0f), 1.0f);
    };
    __stcs(gray + t, make_float4(px(p.x, q.x, k.x), px(p.y, q.y, k.y), px(p.z, q.z, k.z), px(p.w, q.w, k.w)));
}

inline int step_grid(int B) { return (B + WARPS_PER_BLOCK * ENVS_PER_WARP - 1) / (WARPS_PER_BLOCK * ENVS_PER_WARP); }

}  // namespace

extern "C" {

int bk_env_render(int B, const uint64_t *hdr, const uint32_t *bricks, float *state_out, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && bricks && state_out, "bad argument");
    env_render_kernel<<<step_grid(B), WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(B, hdr, bricks, state_out);
    MZB_LAUNCH_CHECK();
    return 0;
}

int bk_env_reset(int B, uint64_t *hdr, uint32_t *bricks, const int64_t *offset, const int64_t *ball_x,
                 const int64_t *ball_h, const int64_t *dx_pick, float *state_out, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && bricks && offset && ball_x && ball_h && dx_pick, "bad argument");
    env_reset_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, hdr, bricks, offset, ball_x, ball_h, dx_pick);
    MZB_LAUNCH_CHECK();
    return state_out ? bk_env_render(B, hdr, bricks, state_out, stream) : 0;
}

int bk_env_reset_device_rng(int B, uint64_t *hdr, uint32_t *bricks, uint64_t seed, uint64_t episode,
                            float *state_out, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && bricks, "bad argument");
    env_reset_rng_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, hdr, bricks, seed, episode);
    MZB_LAUNCH_CHECK();
    return state_out ? bk_env_render(B, hdr, bricks, state_out, stream) : 0;
}

int bk_env_step(int B, uint64_t *hdr, uint32_t *bricks, const int64_t *action, uint8_t *done,
                float *next_state, float *reward, float *valid, float *gray, const float *rewards4,
                int32_t *status, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && bricks && action && done && reward && valid && rewards4 && status, "bad argument");
    const float4 rw = make_float4(rewards4[0], rewards4[1], rewards4[2], rewards4[3]);
    const dim3 grid(step_grid(B)), block(WARPS_PER_BLOCK * 32);
    cudaStream_t st = (cudaStream_t)stream;
#define MZB_STEP(F, G) env_step_kernel<F, G><<<grid, block, 0, st>>>(B, hdr, bricks, action, done, next_state, reward, valid, gray, rw, status)
    if (next_state && gray) MZB_STEP(true, true);
    else if (next_state) MZB_STEP(true, false);
    else if (gray) MZB_STEP(false, true);
    else MZB_STEP(false, false);
#undef MZB_STEP
    MZB_LAUNCH_CHECK();
    return 0;
}

size_t bk_env_io_layout(int B, int want_state, int want_gray, size_t *off8)
{
    // [frames | gray | reward | valid | done | action | status]: everything 16-byte aligned; [done | action | status] is what the host sends
    auto up = [](size_t v) { return (v + 15) & ~(size_t)15; };
    size_t o = 0, off[8];
    off[0] = o; o += want_state ? up((size_t)B * 3840) : 0;
    off[1] = o; o += want_gray ? up((size_t)B * 1280) : 0;
    off[2] = o; o += up((size_t)B * 4);
    off[3] = o; o += up((size_t)B * 12);
    off[4] = o; o += up((size_t)B);
    off[5] = o; o += up((size_t)B * 8);
    off[6] = o; o += 16;
    off[7] = o;
    if (off8) for (int i = 0; i < 8; ++i) off8[i] = off[i];
    return o;
}

int bk_env_step_host(int B, uint64_t *hdr, uint32_t *bricks, void *io_dev, void *host_in, void *host_out, const int64_t *action_host,
                     uint8_t *done_host, int want_state, int want_gray, const float *rewards4, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && bricks && io_dev && host_in && host_out && action_host && done_host && rewards4, "bad argument");
    size_t off[8];
    bk_env_io_layout(B, want_state, want_gray, off);
    uint8_t *io = reinterpret_cast<uint8_t *>(io_dev), *hin = reinterpret_cast<uint8_t *>(host_in);
    const uint8_t *hout = reinterpret_cast<const uint8_t *>(host_out);
    cudaStream_t st = (cudaStream_t)stream;
    static int zc_max = -1;
    if (zc_max < 0) { const char *e = getenv("MZB_ENV_ZEROCOPY_MAX"); zc_max = e ? atoi(e) : 2048; }
    if (B <= zc_max) {
        // Small batches (config.yaml's 24 environments): no copy engine at all.  Pinned host memory is mapped into the device's address
        // space (unified addressing), so the step kernel reads the actions / done flags from the pinned output block and writes frames,
        // rewards, valid-action masks, done flags and the status word straight into it over PCIe: one launch + one synchronisation.
        uint8_t *ho = reinterpret_cast<uint8_t *>(host_out), *dv = nullptr;
        MZB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&dv), ho, 0));
        memcpy(ho + off[4], done_host, (size_t)B);
        memcpy(ho + off[5], action_host, (size_t)B * 8);
        memset(ho + off[6], 0, 16);
        int rc = bk_env_step(B, hdr, bricks, reinterpret_cast<const int64_t *>(dv + off[5]), dv + off[4],
                             want_state ? reinterpret_cast<float *>(dv + off[0]) : nullptr, reinterpret_cast<float *>(dv + off[2]),
                             reinterpret_cast<float *>(dv + off[3]), want_gray ? reinterpret_cast<float *>(dv + off[1]) : nullptr, rewards4,
                             reinterpret_cast<int32_t *>(dv + off[6]), stream);
        if (rc) return rc;
        MZB_CUDA(cudaStreamSynchronize(st));
        memcpy(done_host, ho + off[4], (size_t)B);
        return *reinterpret_cast<const int32_t *>(ho + off[6]) & 0x7fffffff;
    }
    memcpy(hin, done_host, (size_t)B);                                           // the caller's tensors -> the pinned staging block
    memcpy(hin + (off[5] - off[4]), action_host, (size_t)B * 8);
    memset(hin + (off[6] - off[4]), 0, 16);
    MZB_CUDA(cudaMemcpyAsync(io + off[4], hin, off[7] - off[4], cudaMemcpyHostToDevice, st));           // done | action | status = 0
    int rc = bk_env_step(B, hdr, bricks, reinterpret_cast<const int64_t *>(io + off[5]), io + off[4],
                         want_state ? reinterpret_cast<float *>(io + off[0]) : nullptr, reinterpret_cast<float *>(io + off[2]),
                         reinterpret_cast<float *>(io + off[3]), want_gray ? reinterpret_cast<float *>(io + off[1]) : nullptr, rewards4,
                         reinterpret_cast<int32_t *>(io + off[6]), stream);
    if (rc) return rc;
    MZB_CUDA(cudaMemcpyAsync(host_out, io, off[7], cudaMemcpyDeviceToHost, st));
    MZB_CUDA(cudaStreamSynchronize(st));
    memcpy(done_host, hout + off[4], (size_t)B);                                 // done_mask is updated in place (parallel_breakout.py:204,247)
    return *reinterpret_cast<const int32_t *>(hout + off[6]) & 0x7fffffff;
}

int bk_env_ingest(int B, const float *state, const int64_t *ball_dx, const float *ball_dy, uint64_t *hdr,
                  uint32_t *bricks, int32_t *status, void *stream)
{
    MZB_CHECK_ARG(B > 0 && state && hdr && bricks && status, "bad argument");
    env_ingest_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, state, ball_dx, ball_dy, hdr, bricks, status);
    MZB_LAUNCH_CHECK();
    return 0;
}

int bk_env_velocity(int B, const uint64_t *hdr, int64_t *ball_dx, float *ball_dy, void *stream)
{
    MZB_CHECK_ARG(B > 0 && hdr && ball_dx && ball_dy, "bad argument");
    env_velocity_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, hdr, ball_dx, ball_dy);
    MZB_LAUNCH_CHECK();
    return 0;
}

int bk_gray(int B, const float *state, float *gray, void *stream)
{
    MZB_CHECK_ARG(B > 0 && state && gray, "bad argument");
    const size_t n4 = (size_t)B * GRAY_V4;
    gray_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n4, reinterpret_cast<const float4 *>(state),
                                                                              reinterpret_cast<float4 *>(gray));
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
