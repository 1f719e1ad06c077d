/*
 * oracle/breakout_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, per-environment scalar code on the reference's own dense
 * float32 (B,3,16,20) state layout) of the reference's vectorised Breakout environment:
 *     /root/reference/environment/parallel_breakout.py
 *         BreakoutEnvironment.reset              :107-139
 *         BreakoutEnvironment.get_valid_actions  :141-155
 *         BreakoutEnvironment.step               :158-254
 *     /root/reference/train_torch.py
 *         RLSystem.convert_to_grayscale          :334-358
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product path (the .cu files under muzero-breakout_b200/csrc) never links or calls it.
 *
 * Parity pin: the reference has no tests or golden vectors of its own, so this
 * oracle is pinned against outputs of the reference itself, generated in the build container by
 * tests/golden/gen_golden.py (which imports /root/reference) and committed under
 * tests/golden/env_*.npz.  tests/test_oracle_env.py replays them through this file bit-exactly.
 *
 * It deliberately works on the dense planes (argmax of the paddle row, scan for the ball pixel,
 * indexed brick cells) exactly like the reference does, so it shares no design with the CUDA
 * structure-of-arrays kernel it checks.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BK_H 16
#define BK_W 20
#define BK_PADDLE_W 6
#define BK_BRICK_ROWS 3
#define CH_PADDLE 0
#define CH_BALL 1
#define CH_BRICKS 2
#define PLANE (BK_H * BK_W)
#define FRAME (3 * PLANE)

/* reset(): parallel_breakout.py:107-139.  The four random draws are made by the caller (host
 * torch RNG, in the reference's order :116,:126,:127,:136) and passed in:
 *   offset   in [-6, 8)   -> paddle_pos = 20/2 - 6/2 + offset           (:116,:120)
 *   ball_x   in [1, 19)                                                  (:126)
 *   ball_h   in [-3, -1)  -> negative row index, i.e. row 16 + ball_h    (:127-128)
 *   dx_pick  in {0,1}     -> ball_dx = {-1,+1}[dx_pick]                  (:134-136)
 */
void bko_reset(int B, const int64_t *offset, const int64_t *ball_x, const int64_t *ball_h,
               const int64_t *dx_pick, float *state, int64_t *ball_dx, float *ball_dy)
{
    for (int b = 0; b < B; ++b) {
        float *s = state + (size_t)b * FRAME;
        memset(s, 0, sizeof(float) * FRAME);
        int64_t ppos = BK_W / 2 - BK_PADDLE_W / 2 + offset[b];
        for (int k = 0; k < BK_PADDLE_W; ++k)                       /* :123 */
            s[CH_PADDLE * PLANE + (BK_H - 1) * BK_W + (int)(ppos + k)] = 1.0f;
        int64_t row = ball_h[b] < 0 ? BK_H + ball_h[b] : ball_h[b]; /* negative index wraps, :128 */
        s[CH_BALL * PLANE + (int)row * BK_W + (int)ball_x[b]] = 1.0f;
        for (int y = 0; y < BK_BRICK_ROWS; ++y)                     /* :131 */
            for (int x = 0; x < BK_W; ++x) s[CH_BRICKS * PLANE + y * BK_W + x] = 1.0f;
        ball_dx[b] = dx_pick[b] ? 1 : -1;                           /* :134-136 */
        ball_dy[b] = -1.0f;                                         /* :137 */
    }
}

static inline int wrap_row(int y) { return y < 0 ? y + BK_H : y; } /* python negative indexing */

/* step(): parallel_breakout.py:158-254, one environment at a time.
 * rewards = {paddle_hit, brick_hit, game_lost, game_won} (cfg, :82-85).
 * done is read AND written (the reference mutates done_mask in place, :204,:247).
 * Returns 0, or -(b+1) if environment b would raise in the reference (ball leaves the grid /
 * no ball pixel), in which case outputs for that environment are unspecified.
 */
static int step_range(int b0, int b1, const float *state, const int64_t *action, uint8_t *done,
                      int64_t *ball_dx, float *ball_dy, float *next_state, float *reward, float *valid,
                      const float *rewards)
{
    const float r_paddle = rewards[0], r_brick = rewards[1], r_lost = rewards[2], r_won = rewards[3];
    int rc = 0;
    for (int b = b0; b < b1; ++b) {
        const float *s = state + (size_t)b * FRAME;
        float *n = next_state + (size_t)b * FRAME;
        memcpy(n, s, sizeof(float) * FRAME);                        /* clone :173 */
        float r = 0.0f;                                             /* :174 */

        /* paddle :177-186 -- argmax returns the first maximal column */
        const float *prow = s + CH_PADDLE * PLANE + (BK_H - 1) * BK_W;
        int ppos = 0;
        for (int x = 1; x < BK_W; ++x) if (prow[x] > prow[ppos]) ppos = x;
        int a = (int)action[b];
        int pnew = ppos + (a == 0 ? -1 : (a == 2 ? 1 : 0));
        if (pnew < 0) pnew = 0;
        if (pnew > BK_W - BK_PADDLE_W) pnew = BK_W - BK_PADDLE_W;
        float *nprow = n + CH_PADDLE * PLANE + (BK_H - 1) * BK_W;
        for (int x = 0; x < BK_W; ++x) nprow[x] = 0.0f;
        for (int k = 0; k < BK_PADDLE_W; ++k) nprow[pnew + k] = 1.0f;

        /* ball position :189-191 (exactly one ball pixel per env is a reference invariant) */
        int by = -1, bx = -1;
        for (int i = 0; i < PLANE && by < 0; ++i)
            if (s[CH_BALL * PLANE + i] == 1.0f) { by = i / BK_W; bx = i % BK_W; }
        if (by < 0) {
            if (rc == 0) rc = -(b + 1);
            continue;
        }
        float fby = (float)by, fbx = (float)bx;
        int64_t dx = ball_dx[b];
        float dy = ball_dy[b];

        /* wall :195-196, move :198-199 */
        if (fbx + (float)dx < 0.0f || fbx + (float)dx >= (float)BK_W) dx = -dx;
        float ny = fby + dy;
        float nx = fbx + (float)dx;

        /* lost :202-209 */
        int missed = ny >= (float)BK_H;
        if (missed) r = r_lost;
        int d = done[b] | missed;
        if (d) {
            memset(n + CH_BRICKS * PLANE, 0, sizeof(float) * PLANE);
            memset(n + CH_PADDLE * PLANE, 0, sizeof(float) * PLANE);
            dx = 0;
            dy = 0.0f;
        }
        if (missed) ny = 0.0f;

        /* ceiling :213-214 */
        if (ny < 0.0f) { dy = dy * -1.0f; ny = fby; }

        /* bricks :217-226 */
        float old_dy = dy;
        int inx = (int)nx, iny = (int)ny;
        if (inx < 0 || inx >= BK_W || iny >= BK_H) {
            if (rc == 0) rc = -(b + 1);
            continue;
        }
        int cx = inx - (inx % 2);
        float *bricks = n + CH_BRICKS * PLANE;
        int hit = bricks[wrap_row(iny) * BK_W + cx] == 1.0f;
        if (hit) dy = -old_dy;
        bricks[wrap_row(iny) * BK_W + cx] = 0.0f;
        bricks[wrap_row(iny) * BK_W + cx + 1] = 0.0f;
        if (hit) { ny = fby - old_dy; r += r_brick; }
        iny = (int)ny;
        if (iny >= BK_H) {                                          /* reference: IndexError at :243 */
            if (rc == 0) rc = -(b + 1);
            continue;
        }

        /* paddle :229-239 (paddle_mask is built from paddle_positions, also for done envs) */
        int row_hit = ny == (float)(BK_H - 1);
        int on_paddle = inx >= pnew && inx < pnew + BK_PADDLE_W;
        if (row_hit && on_paddle) { dy = -dy; r += r_paddle; }

        /* redraw ball :242-243 (row -1 wraps to row 15) */
        memset(n + CH_BALL * PLANE, 0, sizeof(float) * PLANE);
        n[CH_BALL * PLANE + wrap_row(iny) * BK_W + inx] = 1.0f;

        /* terminal :246-250 */
        int any_brick = 0;
        for (int i = 0; i < PLANE; ++i) if (bricks[i] != 0.0f) { any_brick = 1; break; }
        int finished = !any_brick;
        d |= finished;
        if (d) {
            memset(n + CH_BRICKS * PLANE, 0, sizeof(float) * PLANE);
            memset(n + CH_PADDLE * PLANE, 0, sizeof(float) * PLANE);
        }
        if (finished ^ missed) r += r_won;

        /* valid actions :141-155,:252 */
        valid[b * 3 + 0] = pnew == 0 ? 0.0f : 1.0f;
        valid[b * 3 + 1] = 1.0f;
        valid[b * 3 + 2] = (pnew + BK_PADDLE_W >= BK_W) ? 0.0f : 1.0f;

        reward[b] = r;
        done[b] = (uint8_t)d;
        ball_dx[b] = dx;
        ball_dy[b] = dy;
    }
    return rc;
}

int bko_step(int B, const float *state, const int64_t *action, uint8_t *done, int64_t *ball_dx,
             float *ball_dy, float *next_state, float *reward, float *valid, const float *rewards)
{
    return step_range(0, B, state, action, done, ball_dx, ball_dy, next_state, reward, valid, rewards);
}

/* Same computation split over `nthreads` host threads by contiguous env ranges (envs are
 * independent); used only by bench.py's CPU-baseline legs so they can use every host core. */
typedef struct {
    int b0, b1, rc;
    const float *state; const int64_t *action; uint8_t *done; int64_t *dx; float *dy;
    float *next_state, *reward, *valid; const float *rewards;
} StepJob;

static void *step_job(void *p)
{
    StepJob *j = (StepJob *)p;
    j->rc = step_range(j->b0, j->b1, j->state, j->action, j->done, j->dx, j->dy, j->next_state,
                       j->reward, j->valid, j->rewards);
    return 0;
}

int bko_step_mt(int nthreads, int B, const float *state, const int64_t *action, uint8_t *done,
                int64_t *ball_dx, float *ball_dy, float *next_state, float *reward, float *valid,
                const float *rewards)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    StepJob *jobs = (StepJob *)malloc(sizeof(StepJob) * nthreads);
    int rc = 0;
    for (int i = 0; i < nthreads; ++i) {
        StepJob j = {(int)((int64_t)B * i / nthreads), (int)((int64_t)B * (i + 1) / nthreads), 0,
                     state, action, done, ball_dx, ball_dy, next_state, reward, valid, rewards};
        jobs[i] = j;
        pthread_create(&th[i], 0, step_job, &jobs[i]);
    }
    for (int i = 0; i < nthreads; ++i) {
        pthread_join(th[i], 0);
        if (rc == 0) rc = jobs[i].rc;
    }
    free(th); free(jobs);
    return rc;
}

/* convert_to_grayscale(): train_torch.py:334-358.
 * gray = clamp((paddle*0.3 + ball*1.0) + bricks*0.6, 0, 1), every step rounded to float32. */
void bko_gray(int B, const float *state, float *gray)
{
    for (int b = 0; b < B; ++b) {
        const float *s = state + (size_t)b * FRAME;
        float *g = gray + (size_t)b * PLANE;
        for (int i = 0; i < PLANE; ++i) {
            volatile float p = s[CH_PADDLE * PLANE + i] * 0.3f;
            volatile float q = s[CH_BALL * PLANE + i] * 1.0f;
            volatile float k = s[CH_BRICKS * PLANE + i] * 0.6f;
            volatile float t = p + q;
            float v = t + k;
            g[i] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        }
    }
}
