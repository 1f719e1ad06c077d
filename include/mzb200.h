/*
 * mzb200.h -- C ABI of libmzb200.so: the B200 (sm_100a) acting hot path of MuZero-Breakout.
 *
 * The reference (ulrikisdahl/MuZero-Breakout) has no FFI: its plug points are Python classes loaded
 * by name (utils.py:84-96 get_class; train_torch.py:90-94).  This header is the boundary a binding
 * for that path attaches to; muzero-breakout_b200/{environment/parallel_breakout.py,src/mcts.py}
 * are the ctypes hosts that mirror the reference classes on top of it.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; mzb_last_error() gives the message of the
 *     calling thread's last failure;
 *   - all data pointers are DEVICE pointers owned by the caller (allocated with any CUDA allocator);
 *     nothing is allocated, freed or synchronised behind the caller's back;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden sync;
 *   - no torch (or any C++) type appears in a signature.
 */
#ifndef MZB200_H
#define MZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZB_VERSION 1

int mzb_version(void);
const char *mzb_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t mzb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Breakout environment  (reference: environment/parallel_breakout.py)
 *
 * Authoritative state is structure-of-arrays in HBM:
 *   hdr    uint64[B]      bit 0-4 ball x | 5-8 ball y | 9-12 paddle x | 13 paddle/bricks drawn
 *                         (= not done) | 14-15 ball_dx+1 | 16-17 ball_dy+1 | 32-47 mask of non-empty
 *                         brick rows
 *   bricks uint32[16][B]  row-major over rows; bit x of bricks[y][b] = brick cell (y,x), 20 bits used
 * Frames are the reference's dense float32 (B,3,16,20) planes PADDLE=0, BALL=1, BRICKS=2
 * (parallel_breakout.py:88-90,105).
 * status int32[1]: sticky error bits OR-ed by kernels (never cleared by the library):
 */
#define MZB_ENV_ERR_BALL_LEFT_GRID 1 /* reference would raise IndexError (:243) */
#define MZB_ENV_ERR_BAD_STATE 2      /* ingest: not 0/1, not exactly one ball, malformed paddle row */
#define MZB_ENV_ERR_BAD_ACTION 4     /* action outside {0,1,2} (treated as "stay", like the reference's where()) */

#define MZB_ENV_H 16
#define MZB_ENV_W 20
#define MZB_ENV_FRAME_FLOATS 960 /* 3*16*20 */

/* reset(): parallel_breakout.py:107-139.  The four int64[B] draw arrays are the reference's four
 * RNG calls (:116 offset in [-6,8), :126 ball x in [1,19), :127 ball row offset in [-3,-1),
 * :136 pick in {0,1} -> dx = -1/+1), made by the host in the reference's order.  Writes the SoA state
 * and, if state_out != NULL, the dense frame. */
int bk_env_reset(int B, uint64_t *hdr, uint32_t *bricks, const int64_t *offset, const int64_t *ball_x,
                 const int64_t *ball_h, const int64_t *dx_pick, float *state_out, void *stream);

/* same, with the four draws made on the device from a counter-based generator keyed by
 * (seed, episode, env) -- the throughput mode, no host RNG in the loop. */
int bk_env_reset_device_rng(int B, uint64_t *hdr, uint32_t *bricks, uint64_t seed, uint64_t episode,
                            float *state_out, void *stream);

/* step(): parallel_breakout.py:158-254 + get_valid_actions :141-155.
 *   action int64[B] in {0=left,1=stay,2=right}; done uint8[B] read AND written (the reference mutates
 *   done_mask in place, :204,:247); rewards4 = {paddle_hit, brick_hit, game_lost, game_won} (:82-85).
 *   Outputs: next_state float32[B*960] (may be NULL: SoA only), reward float32[B], valid float32[B*3],
 *   gray float32[B*320] (may be NULL; = convert_to_grayscale(next_state), train_torch.py:334-358). */
int bk_env_step(int B, uint64_t *hdr, uint32_t *bricks, const int64_t *action, uint8_t *done,
                float *next_state, float *reward, float *valid, float *gray, const float *rewards4,
                int32_t *status, void *stream);

/* dense frame -> SoA (used when the caller hands step() a state tensor this library did not
 * produce).  ball_dx int64[B] / ball_dy float32[B] may be NULL (= keep the velocities in hdr). */
int bk_env_ingest(int B, const float *state, const int64_t *ball_dx, const float *ball_dy, uint64_t *hdr,
                  uint32_t *bricks, int32_t *status, void *stream);

/* SoA -> dense frame / velocities (the reference's .ball_dx int64, .ball_dy float32 attributes) */
int bk_env_render(int B, const uint64_t *hdr, const uint32_t *bricks, float *state_out, void *stream);
int bk_env_velocity(int B, const uint64_t *hdr, int64_t *ball_dx, float *ball_dy, void *stream);

/* convert_to_grayscale(): train_torch.py:334-358 on an arbitrary dense state */
int bk_gray(int B, const float *state, float *gray, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MZB200_H */
