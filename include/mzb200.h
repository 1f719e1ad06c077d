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
/* ABI check for foreign-language bindings: sizeof(mz_tree_args) (which=0), sizeof(mz_op) (which=1), sizeof(rb_ring) (2) */
size_t mzb_sizeof(int which);

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

/* step() for HOST tensors (the reference's call: CPU action / done_mask in, CPU tensors out, parallel_breakout.py:158-254) as one call:
 * ONE host->device copy, the step kernel, ONE device->host copy, one stream synchronisation.
 *   bk_env_io_layout  byte offsets of {frames, gray, reward, valid, done, action, status, total} in the packed buffer (16-byte aligned
 *                     regions; frames / gray have size 0 when not wanted); returns the total size.
 *   bk_env_step_host  io_dev: device buffer of `total` bytes; host_in: pinned staging block of total - offset[done] bytes (the
 *                     [done | action | status] part of the layout; filled by the call); host_out: pinned, `total` bytes, receives
 *                     the whole packed buffer; action_host int64[B] / done_host uint8[B]: the caller's host arrays (done_host is
 *                     read AND written, like the reference's in-place done_mask).  Returns the status bits of this step
 *                     (MZB_ENV_ERR_*, >= 0) or a negative error code.  Up to MZB_ENV_ZEROCOPY_MAX environments (default 2048) the
 *                     copy engine is not used at all: pinned host memory is mapped into the device's address space, the kernel reads
 *                     actions / done flags from host_out's regions and writes every output straight into it (host_in is unused). */
size_t bk_env_io_layout(int B, int want_state, int want_gray, size_t *off8);
int bk_env_step_host(int B, uint64_t *hdr, uint32_t *bricks, void *io_dev, void *host_in, void *host_out, const int64_t *action_host,
                     uint8_t *done_host, int want_state, int want_gray, const float *rewards4, void *stream);

/* dense frame -> SoA (used when the caller hands step() a state tensor this library did not
 * produce).  ball_dx int64[B] / ball_dy float32[B] may be NULL (= keep the velocities in hdr). */
int bk_env_ingest(int B, const float *state, const int64_t *ball_dx, const float *ball_dy, uint64_t *hdr,
                  uint32_t *bricks, int32_t *status, void *stream);

/* SoA -> dense frame / velocities (the reference's .ball_dx int64, .ball_dy float32 attributes) */
int bk_env_render(int B, const uint64_t *hdr, const uint32_t *bricks, float *state_out, void *stream);
int bk_env_velocity(int B, const uint64_t *hdr, int64_t *ball_dx, float *ball_dy, void *stream);

/* convert_to_grayscale(): train_torch.py:334-358 on an arbitrary dense state */
int bk_gray(int B, const float *state, float *gray, void *stream);


/* ------------------------------------------------------------------------------------------------
 * Latent MCTS tree bookkeeping  (reference: src/mcts.py)
 *
 * One flat, preallocated block per root in HBM (mz_tree_bytes(num_simulations) bytes each, 16-byte
 * rows: Q|vsum, P, R, N|child, then selection meta + path); num_simulations + 2 node slots per tree,
 * slot 0 = root.  Latents of expanded nodes live in a caller-owned store
 * latent_store[tree][slot][latent_bytes].
 *
 * A search is: root prediction -> mz_tree_root -> for sim in 0..S-1 { dynamics+prediction on the
 * selected leaves -> mz_tree_step(sim) }.  mz_tree_step(sim) = _backup of simulation `sim`
 * (mcts.py:203-234) fused with _select_nodes of simulation sim+1 (:136-182) and the gather of the
 * selected parent latents into dyn_in; on the last simulation it writes _compute_results (:236-250)
 * instead.  Inputs value/reward are ALREADY inverse-transformed scalars
 * (ScalarTransforms.inverted_softmax_expectation, utils.py:74-81) and pi are softmax probabilities
 * (mcts.py:100,199), float32.
 */
typedef struct mz_tree_args {
    int32_t B;               /* number of roots */
    int32_t num_simulations; /* cfg["num_simulations"] (mcts.py:13) */
    int32_t sim;             /* mz_tree_step: index of the simulation being backed up */
    int32_t reserved;
    void *trees;             /* [B] blocks of mz_tree_bytes() */
    const float *s_tab;      /* [S+1] float32(sqrt(n))                       } from mz_puct_tables, */
    const float *k_tab;      /* [S+1] float32(c1 + log((n + c2 + 1) / c2))   } uploaded by the caller */
    double discount;         /* cfg["search"]["discount_factor"] (mcts.py:17) */
    double noise_weight;     /* MCTSSearchVec.noise_weight (mcts.py:22), root only */
    uint64_t seed;           /* tie-break stream key */
    const float *reward;     /* [B]    leaf edge rewards (step) */
    const float *value;      /* [B]    leaf values (step) / root values (root) */
    const float *pi;         /* [B][3] leaf priors (step) / root priors before noise (root) */
    const float *noise;      /* [B][3] Dirichlet(0.25) samples (root; mcts.py:114) */
    int32_t *leaf_parent;    /* [B] out: slot of the selected leaf's parent */
    int32_t *leaf_action;    /* [B] out: action on that edge */
    int32_t *leaf_slot;      /* [B] out: slot the new leaf will occupy */
    const void *latent_store; /* may be NULL together with dyn_in */
    void *dyn_in;            /* [B][latent_bytes] out: parent latents, input of the dynamics network */
    int64_t latent_bytes;    /* multiple of 16 */
    float *out_value;        /* [B]    root value (last step) */
    int64_t *out_visits;     /* [B][3] root visit counts (last step) */
    int32_t *depth_hist;     /* optional [S+1] histogram of selection depths, or NULL */
    const uint64_t *seed_dev; /* optional device word: if not NULL it replaces `seed` (lets a captured CUDA graph
                                 be replayed with a new stream key) */
} mz_tree_args;

size_t mz_tree_bytes(int num_simulations);
int mz_tree_nodes(int num_simulations);
/* host-side: the two pUCT terms the reference computes in Python doubles (mcts.py:286-289) */
int mz_puct_tables(int num_simulations, double c1, double c2, float *s_tab_host, float *k_tab_host);
int mz_tree_root(const mz_tree_args *args, void *stream);
int mz_tree_step(const mz_tree_args *args, void *stream);


/* ------------------------------------------------------------------------------------------------
 * MuZero networks  (reference: src/networks.py, utils.py:74-81)
 *
 * Activations are channels-last: [sample][y][x][channel], float32 (exact path, CUDA-core FFMA,
 * 1e-5 parity mode) or bfloat16 (tcgen05 tensor-core path, fp32 accumulation in TMEM).  A network
 * evaluation is a short program of mz_op records run in order by mz_run() on one stream; the host
 * builds the programs once per (weights, batch size) from MuZeroAgent.state_dict()
 * (muzero-breakout_b200/src/networks.py) with every pointer already resolved.
 *
 *   MZ_OP_CONV      dst = act((conv_k(src) [+ act_bias[act_idx]]) * scale + shift [+ res])
 *                   conv_k: k x k, stride 1, zero padding k/2 (nn.Conv2d of networks.py:12,26,28,47,65);
 *                   scale/shift = conv bias + eval-mode BatchNorm folded per channel (:16-17,31-35); scale == NULL
 *                   means 1 (the BatchNorm scale is already folded into w: what the 16-bit packers do);
 *                   16-bit residual stream (res_lo / dst_lo, optional): the value added is res + res_lo / LO_SCALE and
 *                   the rounding error of the 16-bit dst is kept as dst_lo = e4m3((x - dst) * LO_SCALE), LO_SCALE = 2048
 *                   (fp16) / 256 (bf16) -- the stream of a ResidualBlock chain (:31-35) then carries 15 / 12
 *                   significant bits instead of 11 / 8 while the tensor-core operands stay 16-bit.  A correction
 *                   plane is mz_conv_lo_bytes() bytes in a layout private to the convolution kernels (it is only
 *                   ever exchanged between convolutions of the same shape and batch);
 *                   act_bias = contribution of the three spatially-constant one-hot action planes of
 *                   the dynamics input (:295), a [3][H*W][cout] table, selected per sample by act_idx
 *   MZ_OP_POOL2     2x2 average pool, stride 2 (:43,82,92)
 *   MZ_OP_SCALE     MuZeroAgent._scale_state (:314-328): per sample (x - min) / (max - min + 1e-8) over
 *                   all H*W*C elements of the fp32 src; written to dst and, if dst2 != NULL, also to
 *                   dst2 + (i * dst2_stride + dst2_slot[i]) * (H*W*C*elsize)   (the tree's latent store)
 *   MZ_OP_HEAD      Flatten(C,H,W) -> Linear (:147-149,207-209,221-223) on a channels-last src (cin channels read out of
 *                   pixel rows of cout channels; cout = 0 means cin: a dense tensor), then
 *                   head_mode 0: raw logits; 1: inverted_softmax_expectation (utils.py:74-81) -> out[n];
 *                   2: softmax probabilities -> out[n][nout] (mcts.py:100,199)
 *   MZ_OP_NCHW_IN   float32 NCHW src -> channels-last dst (+ optional dst2 as for MZ_OP_SCALE)
 *   MZ_OP_NHWC_OUT  channels-last src -> float32 NCHW dst
 */
enum { MZ_OP_CONV = 0, MZ_OP_POOL2 = 1, MZ_OP_SCALE = 2, MZ_OP_HEAD = 3, MZ_OP_NCHW_IN = 4, MZ_OP_NHWC_OUT = 5 };
enum { MZ_F32 = 0, MZ_BF16 = 1, MZ_F16 = 2 };   /* MZ_F16: same tensor-core path and rate as bf16, 11-bit mantissa */
enum { MZ_ACT_NONE = 0, MZ_ACT_RELU = 1, MZ_ACT_LEAKY_RELU = 2, MZ_ACT_SILU = 3, MZ_ACT_GELU = 4 }; /* utils.py:99-108 */

typedef struct mz_op {
    int32_t op;        /* MZ_OP_* */
    int32_t dtype;     /* MZ_F32 / MZ_BF16 / MZ_F16: element type of src / dst / res / w */
    int32_t H, W;      /* spatial size of src */
    int32_t cin, cout; /* channels of src / dst */
    int32_t ksize;     /* conv: 1 or 3 */
    int32_t act;       /* MZ_ACT_* */
    int32_t use_tc;    /* conv, bf16: 1 = tcgen05 kernel, 0 = CUDA-core kernel */
    int32_t nout;      /* head: output features (<= 16) */
    int32_t head_mode; /* head: 0 logits, 1 scalar, 2 probabilities */
    int32_t w_layout;  /* conv: 0 = w[cout][ksize*ksize*cin]; 1 (tensor-core path only) = tile-contiguous
                          w[tap][cin/64][cout][64], each 64-channel weight tile stored as consecutive 128-byte rows */
    const void *src;
    void *dst;
    const void *res;         /* conv: residual (same shape as dst) or NULL */
    float *dst_f32;          /* conv / pool: optional extra float32 copy of dst (input of MZ_OP_SCALE) */
    const void *w;           /* conv: [cout][ksize*ksize*cin], tap-major then channel; head: float32 [nout][H*W*cin] in (y,x,c) order */
    const float *scale;      /* [cout], or NULL = 1 */
    const float *shift;      /* [cout]; head: bias [nout] */
    const float *act_bias;   /* conv: [3][H*W][cout] or NULL */
    const int32_t *act_idx;  /* conv: [n] */
    void *dst2;              /* scale / nchw_in: second destination base or NULL */
    const int32_t *dst2_slot; /* [n] or NULL (= slot 0) */
    int64_t dst2_stride;     /* slots per sample in dst2 */
    float *out;              /* head: see head_mode */
    float *out_logits;       /* head: optional raw logits [n][nout] */
    const void *res_lo;      /* conv, 16-bit: correction plane of res or NULL (= 0) */
    void *dst_lo;            /* conv, 16-bit: correction plane of dst or NULL (not kept) */
    const float *res_f32;    /* conv, 16-bit: the residual as float32 [n][H][W][cout] INSTEAD of res / res_lo (a stream that enters
                                the network as float32, e.g. the latent argument of MuZeroAgent.evaluate_state), or NULL */
    double *bn_partial;      /* conv, training form only (dst == NULL, dst_f32 set, act none, no residual but res_f32): the kernel also
                                writes the per-channel sums and sums of squares of its float32 output over every 32-row group of its tiles,
                                [mz_conv_stats_blocks()][2][cout] doubles -- the partial sums of the training-mode BatchNorm that follows
                                (mz_bn_train_fwd_pre), so that no separate reduction pass reads the output again.  NULL: not computed */
} mz_op;

int mz_run(const mz_op *ops, int n_ops, int nsamples, void *stream);
/* number of partial-sum blocks a tensor-core convolution over nsamples images of H x W writes to mz_op.bn_partial */
int mz_conv_stats_blocks(int nsamples, int H, int W, int ksize);
/* bytes of a correction plane (res_lo / dst_lo) of an [nsamples][H][W][cout] convolution output */
size_t mz_conv_lo_bytes(int nsamples, int H, int W, int cout, int ksize);


/* ------------------------------------------------------------------------------------------------
 * Acting-move glue  (reference: train_torch.py _prepare_mcts_input :259-277, _encode_actions :279-293,
 * _pad_initial_state :313-332, temperature sampling :192-198)
 *
 * mz_rep_input: observation history -> representation-network input, channels-last [B][16*20][64]:
 *   channels 0..30 = the last 31 gray frames appended to the trajectory (oldest first), 31 = the current
 *   frame, 32..63 = the last 32 actions / 3 as constant planes (oldest first).
 *   frames float32 [slots][B][320] ring with `head` = slot of the newest appended frame; cur float32 [B][320];
 *   acts int32 [slots][B] ring with `ahead` = slot of the newest action; slots >= 32; dtype MZ_F32 / MZ_BF16.
 * mz_sample_actions: p_a = visits_a ** (1/temperature) / sum, one categorical draw per env with
 *   u = u32(seed, env, step) / 2^32; writes action int64 [B] (+ optional int32 copy into an action-ring slot,
 *   optional probabilities float32 [B][3]).
 */
int mz_rep_input(int B, int slots, const float *frames, int head, const float *cur, const int32_t *acts, int ahead, void *out,
                 int dtype, void *stream);
int mz_sample_actions(int B, const int64_t *visits, double temperature, uint64_t seed, uint32_t step, int64_t *action,
                      int32_t *act_slot, float *probs_out, void *stream);


/* ------------------------------------------------------------------------------------------------
 * Fused residual trunk: a run of consecutive MZ_OP_CONV records that are all 16-bit / use_tc / w_layout 1 /
 * scale == NULL / 256 -> 256 on the 4x5 latent (3x3; 1x1 only as trailing records whose outputs nothing in the run
 * reads: the head ConvBlocks, networks.py:138-146,200-218) and touch at most MZ_STACK_MAX_BUFS activation buffers is
 * executed by ONE persistent launch (csrc/conv_stack.cu); layers are ordered per 128-sample group through device
 * counters instead of kernel boundaries.
 *   mz_stack_layer_bytes()  size of one device-resident layer descriptor
 *   mz_stack_build          fills a HOST blob (64-byte aligned, n_ops * mz_stack_layer_bytes() bytes) from the op
 *                           records; bufs[n_bufs] are the activation buffers the ops' src/dst/res point to.
 *                           The caller copies the blob to device memory once.
 *   mz_stack_scratch_bytes  size of the dependency-counter scratch of a launch over nsamples samples
 *   mz_stack_run            runs the trunk on samples [sample0, sample0 + nsamples) of the buffers (sample0 a multiple of
 *                           256): blob_dev = the uploaded blob, bufs = the same buffers in the same order (base
 *                           pointers), act_idx as in mz_op (base pointer), scratch = mz_stack_scratch_bytes(n_layers,
 *                           nsamples) bytes, 16-byte aligned, ZEROED ONCE by the caller and then owned by the launches
 *                           of this (trunk, nsamples): they keep a launch epoch in it, so a CUDA-graph replay needs no
 *                           reset.  slice_samples (a multiple of 256, or 0): the launch walks the samples slice by slice,
 *                           all layers of a slice before the next one, so that the live activations of a slice (hosts:
 *                           2048 samples = 52 MB) stay in the L2 while the pipelines never drain between slices.
 */
#define MZ_STACK_MAX_BUFS 5
size_t mz_stack_layer_bytes(void);
size_t mz_stack_scratch_bytes(int n_layers, int nsamples);
int mz_stack_build(const mz_op *ops, int n_ops, void *blob_host, size_t blob_bytes, const void *const *bufs, int n_bufs);
int mz_stack_run(const void *blob_dev, int n_layers, int sample0, int nsamples, int slice_samples, void *const *bufs, int n_bufs,
                 const int32_t *act_idx, void *scratch, int dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Latency-mode residual trunk (csrc/conv_lat.cu): the same run of stackable convolutions as mz_stack_*, for SMALL
 * leaf batches (config.yaml's default acting stage, train_torch.py:164-258 with n_parallel = 24: every network call of
 * src/mcts.py:95,194-198 has 24 samples).  One persistent launch of 16 * ceil(nsamples/3) work items
 * (3-sample row tile x 16-output-channel slice; warp-level mma.sync, weights streamed by TMA one item ahead, layers handed over
 * through flag-in-data words in L2).  Results are the same convolution on the same 16-bit operands with fp32
 * accumulation; the accumulation ORDER differs from the tcgen05 kernels, so outputs agree to fp32 rounding, not bit
 * for bit.
 *   mz_lat_layer_bytes()   size of one device-resident layer descriptor
 *   mz_lat_max_samples()   largest batch hosts run here (108 = four waves of work items; measured crossover with the tcgen05 trunk)
 *   mz_lat_build           fills a HOST blob (64-byte aligned, n_ops * mz_lat_layer_bytes() bytes) from the op records
 *                          (a weight tensor map + operand pointers per record); the caller copies it to device memory once.
 *                          Records: 3x3 or 1x1 256->256 convolutions forming a chain (ops[i].src == ops[i-1].dst; the 1x1 is the
 *                          reward head's ConvBlock, networks.py:138-146), optionally followed by ONE pair of 256->128
 *                          convolutions that both read the last 256-channel output and write different buffers (the policy
 *                          3x3 and value 1x1 ConvBlocks, networks.py:200-218): the pair runs as one split layer, 8 + 8
 *                          channel slices.  Up to two trailing MZ_OP_HEAD / MZ_OP_SCALE records (3 or 11 head outputs; 5120-element
 *                          _scale_state) become tail ops: they run per sample at the end of the same launch.  Returns flags >= 0
 *                          (bit 0: the convolutions end with a split pair; bits 1-2: number of tail ops), < 0 on error.
 *   mz_lat_scratch_bytes   size of the launch's scratch (layer hand-off buffers + epoch / counters) for nsamples
 *   mz_lat_run             runs the n_ops records on samples [0, nsamples): flags = mz_lat_build's return value,
 *                          act_idx as in mz_op, scratch = mz_lat_scratch_bytes(nsamples) bytes, 16-byte aligned, ZEROED ONCE by
 *                          the caller and then owned by this trunk's launches (they keep a launch epoch in it, so a
 *                          CUDA-graph replay needs no reset; one scratch per concurrently running trunk)
 */
size_t mz_lat_layer_bytes(void);
int mz_lat_max_samples(void);
int mz_lat_max_layers(void);    /* layer descriptors of one launch are staged in shared memory: at most this many (32) */
int mz_lat_trace(unsigned long long *host_out_8x64);   /* profiling aid (MZB_LAT_TRACE=1): per-layer phase timestamps of CTA 0 */
int mz_lat_build(const mz_op *ops, int n_ops, void *blob_host, size_t blob_bytes);
size_t mz_lat_scratch_bytes(int nsamples);
int mz_lat_run(const void *blob_dev, int n_ops, int flags, int nsamples, const int32_t *act_idx, void *scratch, int dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Replay buffer  (reference: replay_buffer.py)
 *
 * A trajectory of T moves is stored once, as T+1 consecutive entries of an entry ring (entry 0 = the
 * initial gray frame; entry m+1 = move m: gray frame after the move, action, reward, visit counts,
 * search value -- the tuple train_torch.py:205-207 appends), instead of one materialised 32-frame
 * window per sample (replay_buffer.py:116-119).  A sample is 8 bytes of metadata
 * (ring index of entry 0 | start s << 32 | T << 48) + K value targets + the trajectory's reward sum, in a
 * FIFO sample ring of cap_samples = ReplayBuffer.max_length (:154-163).  state[0] = entries ever written,
 * state[1] = samples ever appended (device-resident: appends need no host sync); live samples are the last
 * min(state[1], cap_samples), logical index 0 = oldest (the reference's list order).
 * gpow[k] = (float)(discount ** k) for k = 0..15, computed by the host in double like the reference's
 * Python `self.discount**k` (:142-148).
 * status int32[1]: sticky error bits OR-ed by kernels: */
#define MZB_RB_ERR_BAD_LENGTH 1 /* a trajectory length outside [0, min(T, max_moves)] */
#define MZB_RB_ERR_BAD_INDEX 2  /* rb_gather index outside [-length, length): the reference raises IndexError */

typedef struct rb_ring {
    int32_t cap_samples, cap_entries, K, hist; /* hist = seq_len = 32 (config.yaml:37) */
    float *frame;                              /* [cap_entries][320] */
    int32_t *action;                           /* [cap_entries] */
    float *reward, *value;                     /* [cap_entries] */
    float *visits;                             /* [cap_entries][3]  (float: the reference's stack promotes to fp32) */
    uint64_t *meta;                            /* [cap_samples] */
    float *reward_sum;                         /* [cap_samples] */
    float *target;                             /* [cap_samples][K]  value targets (:136-152) */
    uint64_t *state;                           /* [2] */
    float gpow[16];
} rb_ring;

/* bytes of the per-call placement scratch of rb_append for B trajectories */
size_t rb_plan_bytes(int B);
/* entries the entry ring needs so that no live sample's trajectory is ever overwritten */
long long rb_entries_for(int cap_samples, int K, int max_moves);

/* save_observation_trajectory (:96-165) for B trajectories at once, recorded move-major by the acting loop:
 * action int64[T][B], reward/value float[T][B], visits int64[T][B][3], frames float[T][B][320] (gray frame after
 * the move), init_frame float[B][320]; pad_action = the action the 32 padding rows hold (0 in
 * _pad_initial_state, train_torch.py:324).  Trajectory b has lengths[b] moves if lengths != NULL, else the number of
 * set bytes in recorded[.][b] (uint8 [T][B], "not done before the move", train_torch.py:205).  Trajectories
 * shorter than min_length are skipped (train_torch.py:224 passes K+2; the raw method corresponds to 0);
 * the rest give length-K+1 samples each, appended in env order, evicting the oldest beyond cap_samples. */
int rb_append(const rb_ring *ring, int B, int T, const int64_t *action, const float *reward, const float *value,
              const int64_t *visits, const float *frames, const float *init_frame, int pad_action,
              const uint8_t *recorded, const int32_t *lengths, int min_length, int max_moves, void *plan,
              int32_t *status, void *stream);

/* get_batched_past_actions / future_actions / states / rewards / visit_counts / values (:167-210) for the n
 * logical sample indices idx (device int64; negative = from the newest, like a Python list); any output may be
 * NULL.  past_actions int64[n][hist], future_actions int64[n][K], states float[n][hist][320],
 * rewards float[n][K], visit_counts float[n][K][3], values float[n][K] (the value TARGETS, :152),
 * value_buffer float[n][K] (the raw search values, :131-133), reward_sums float[n] (:122). */
int rb_gather(const rb_ring *ring, int n, const int64_t *idx, int64_t *past_actions, int64_t *future_actions,
              float *states, float *rewards, float *visit_counts, float *values, float *value_buffer,
              float *reward_sums, int32_t *status, void *stream);

/* The representation-network input of a training minibatch, straight from the rings: out float[n][2*hist][320] =
 * torch.cat((get_batched_states(idx).view(n, hist, 16, 20), _encode_actions(get_batched_past_actions(idx))), dim=1)
 * (train_torch.py:392,500; _encode_actions :279-293 = past_action / n_actions as constant planes). */
int rb_gather_input(const rb_ring *ring, int n, const int64_t *idx, int n_actions, float *out, int32_t *status, void *stream);


/* ------------------------------------------------------------------------------------------------
 * Training step: the loss and the optimizer update here; the forward and backward passes of the
 * three networks further down (weight gradients, training-mode BatchNorm, the small layers: pools, Linear heads, _scale_state,
 * action planes); the data gradient of a convolution is mz_run's MZ_OP_CONV on the transposed, tap-flipped filter (mz_pack_conv).
 *
 * mz_loss replaces loss_fn (train_torch.py:33-66) + ScalarTransforms.supports_representation
 * (utils.py:30-64) + what loss.backward() (train_torch.py:515) produces for the three logit tensors.
 *   rows = batch * K samples (the .view(-1, n) of :49-63); supports float32[n_supports] =
 *   ScalarTransforms.supports (utils.py:19); pred_reward / pred_value float32[rows][n_supports] raw logits,
 *   pred_policy float32[rows][n_actions] raw logits; observed_reward / value_target float32[rows] scalars;
 *   visit_counts float32[rows][n_actions] (un-normalised, :58).
 *   losses float32[4] = {(1/K)(reward + value + policy), reward_loss, value_loss, policy_loss}, each a
 *   kl_div(..., reduction="batchmean") = sum / rows.
 *   d_reward / d_value / d_policy (same shapes as the logits, any may be NULL) = d losses[0] / d logits.
 *   scratch: mz_loss_scratch_bytes(rows) bytes, 8-byte aligned, ZEROED ONCE by the caller (the kernel
 *   leaves it ready for the next call); reduction order is fixed, results are deterministic.
 *
 * mz_adam replaces MuZeroAgent.optimizer.step() (networks.py:268: torch.optim.Adam(lr, weight_decay=1e-4),
 * betas (0.9, 0.999), eps 1e-8) for ONE flat float32 buffer of n parameters (16-byte aligned): L2 weight
 * decay added to the gradient, bias-corrected moments, same operation order as torch's single-tensor path.
 * step = number of updates including this one (torch's state["step"] after the increment).
 */
size_t mz_loss_scratch_bytes(int rows);
int mz_loss(int rows, int K, int n_supports, int n_actions, const float *supports, const float *pred_reward, const float *pred_value,
            const float *pred_policy, const float *observed_reward, const float *value_target, const float *visit_counts,
            float *losses, float *d_reward, float *d_value, float *d_policy, void *scratch, void *stream);
int mz_adam(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, double lr, double beta1, double beta2,
            double eps, double weight_decay, int step, void *stream);
/* The same update with the step count on the device, for a training step replayed as a CUDA graph: `state` is MZ_ADAM_STATE_BYTES of
 * device memory, 16-byte aligned, whose first int32 holds the number of updates done so far (zero it once, or store torch's state["step"]);
 * the call increments it and derives this update's bias corrections from it on the device (two launches, no host value of the step). */
#define MZ_ADAM_STATE_BYTES 64
int mz_adam_dev(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, double lr, double beta1, double beta2,
                double eps, double weight_decay, void *state, void *stream);

/* Weight gradient of a stride-1 "same" 3x3 / 1x1 convolution with 256 input and 256 output channels (the residual trunks,
 * networks.py:24-25; autograd of nn.Conv2d inside loss.backward(), train_torch.py:515) on the tensor cores (csrc/wgrad.cu):
 *   dw float32[256][256][ksize][ksize] (PyTorch's weight layout) = sum_{n,y,x} dy[n][y][x][co] * x[n][y+ky-pad][x+kx-pad][ci].
 * Both operands are given TRANSPOSED to channel-major [256][H*W][ns] (ns = mz_wgrad_padded_samples(n), padding samples zero),
 * 16-bit (dtype MZ_BF16 / MZ_F16): mz_wgrad_transpose makes that layout from a channels-last [n][H*W][C] tensor.
 * partial: mz_wgrad_partial_bytes(ksize, n) bytes of scratch (per-tap, per-K-split fp32 partial sums, reduced in a fixed order). */
int mz_wgrad_padded_samples(int n);
size_t mz_wgrad_partial_bytes(int ksize, int n);
int mz_wgrad_transpose(int n, int P, int C, const void *src, void *dst, void *stream);
/* f16_to_bf16 != 0: src is float16, dst bfloat16 -- the activations of an fp16 forward pass as the bf16 operand next to bf16 gradients
 * (one tcgen05 kind::f16 MMA cannot mix A / B element types) */
int mz_wgrad_transpose_cvt(int n, int P, int C, const void *src, void *dst, int f16_to_bf16, void *stream);
/* the same into the sample columns [s_offset, s_offset + mz_wgrad_padded_samples(n)) of a destination with ns_total sample columns per
 * (channel, pixel) row (both multiples of 64): several activations that share a weight -- the K unroll steps of a training step --
 * concatenated along the GEMM's reduction axis, so that ONE mz_conv_wgrad_any(n = ns_total) call gives the sum of their weight gradients */
int mz_wgrad_transpose_into(int n, int P, int C, const void *src, void *dst, int ns_total, int s_offset, int f16_to_bf16, void *stream);
/* both operands of one weight gradient (a: dy with Ca channels, b: x with Cb channels; cvt_*: float16 -> bfloat16) in ONE launch */
int mz_wgrad_transpose_pair(int n, int P, int Ca, const void *src_a, void *dst_a, int cvt_a, int Cb, const void *src_b, void *dst_b, int cvt_b, int ns_total,
                            int s_offset, void *stream);
int mz_conv_wgrad(int n, int H, int W, int ksize, int dtype, const void *dy_t, const void *x_t, float *partial, float *dw, void *stream);
/* accumulate != 0: dw += the gradient (one rounding per call, in call order): the K unroll steps of a training step share their weights
 * (train_torch.py:507-525) and add straight into the parameter's .grad instead of through K separate add kernels */
int mz_conv_wgrad_accum(int n, int H, int W, int ksize, int dtype, const void *dy_t, const void *x_t, float *partial, float *dw, int accumulate,
                        void *stream);

/* The same for the other convolutions of the three networks (representation stems :47,65 and 128-channel blocks, head ConvBlocks
 * :139,201,213, the 256 hidden-state input channels of the dynamics ConvBlock :117): cout in {128, 256}, cin in {64, 128, 256}.
 * dy_t is [cout][H*W][ns], x_t [cin][H*W][ns]; dw float32[cout][dw_cin][ksize][ksize] with dw_cin >= cin: the gradient of the FIRST cin
 * input channels (the dynamics ConvBlock has 256 + 3 action-plane channels: mz_planes_conv_wgrad fills the rest);
 * partial: mz_wgrad_partial_bytes_any() bytes. */
size_t mz_wgrad_partial_bytes_any(int ksize, int n, int cout, int cin);
int mz_conv_wgrad_any(int n, int H, int W, int ksize, int dtype, int cout, int cin, int dw_cin, const void *dy_t, const void *x_t, float *partial,
                      float *dw, int accumulate, void *stream);

/* Training-mode nn.BatchNorm2d of a ConvBlock / ResidualBlock (networks.py:12,16-17,26-35 under train_mode(), train_torch.py:372)
 * on channels-last rows: z float32[M][C] = the convolution output incl. its bias, M = samples * H * W.
 *   forward:  batch mean / biased variance per channel -> save_mean, save_invstd = 1/sqrt(var + eps) (float32[C]);
 *             y = act(gamma * (z - mean) * invstd + beta (+ res)), written 16-bit (y, dtype MZ_BF16 / MZ_F16) and/or float32 (y_f32);
 *             running_mean / running_var (may be NULL) updated in place with torch's rule (momentum, unbiased variance).
 *   backward: g = dy * act'(pre-activation) (dy float32[M][C]); dbeta = sum g, dgamma = sum g * xhat,
 *             dz = gamma * invstd * (g - dbeta/M - xhat * dgamma/M) as float32 (dz) and/or 16-bit (dz16, the operand of the
 *             convolution gradients); dres (may be NULL) = g, the gradient of the residual input.
 * res: the residual added before the activation (16-bit, NULL for a ConvBlock).  act: MZ_ACT_NONE / RELU / LEAKY_RELU.
 * scratch: mz_bn_scratch_bytes(M, C) bytes (fp64 per-CTA partial sums, added in a fixed order: deterministic). */
size_t mz_bn_scratch_bytes(int M, int C);
int mz_bn_train_fwd(int M, int C, const float *z, const float *gamma, const float *beta, const void *res, int dtype, int act, double eps,
                    double momentum, float *running_mean, float *running_var, float *save_mean, float *save_invstd, void *y, float *y_f32,
                    void *scratch, void *stream);
int mz_bn_train_bwd(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int act,
                    const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dz, void *dz16, float *dres, void *scratch,
                    void *stream);
/* mz_bn_train_fwd with the statistics' partial sums already computed by the producing convolution (mz_op.bn_partial: nblocks =
 * mz_conv_stats_blocks() blocks of [2][C] doubles): finalize + apply only */
int mz_bn_train_fwd_pre(int M, int C, int nblocks, const double *partial, const float *z, const float *gamma, const float *beta, const void *res,
                        int dtype, int act, double eps, double momentum, float *running_mean, float *running_var, float *save_mean,
                        float *save_invstd, void *y, float *y_f32, void *stream);
/* the same with separate element types for res (dtype: what the forward pass produced) and dz16 (dz_dtype: the operand of the convolution
 * gradients): an fp16 forward pass with bf16 gradients */
int mz_bn_train_bwd_mixed(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int dz_dtype,
                          int act, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dz, void *dz16, float *dres,
                          void *scratch, void *stream);
/* the same, and dgamma_acc[c] += dgamma[c], dbeta_acc[c] += dbeta[c] when those pointers are not NULL: the parameters' own .grad tensors
 * (the K unroll steps of a training step share their layers, train_torch.py:507-525) without an add kernel per parameter and call */
int mz_bn_train_bwd_acc(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int dz_dtype,
                        int act, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dgamma_acc, float *dbeta_acc,
                        float *dz, void *dz16, float *dres, void *scratch, void *stream);

/* out[c] (+)= sum over the M rows of x float32[M][C]: the bias gradient of a convolution that no BatchNorm follows (the representation
 * network's stem convolutions, networks.py:47,65).  scratch: mz_bn_scratch_bytes(M, C) bytes.  Deterministic. */
int mz_colsum(int M, int C, const float *x, float *out, int accumulate, void *scratch, void *stream);

/* ------------------------------------------------------------------------------------------------
 * The small layers of a training step around the convolutions (csrc/train_layers.cu), forward and backward, on channels-last
 * float32 tensors [n][H][W][C].  All deterministic (fixed-order reductions).
 *
 * mz_cvt16             n float32 values (n % 4 == 0) -> 16-bit (dtype MZ_BF16 / MZ_F16): the operand of the next convolution
 * mz_pool2_train_fwd   nn.AvgPool2d(2, 2) (networks.py:43,82,92): x [n][H][W][C] -> y float32 and / or y16 [n][H/2][W/2][C]
 * mz_pool2_train_bwd   dx [n][H][W][C] = dy [n][H/2][W/2][C] / 4
 * mz_linear_fwd        nn.Flatten + nn.Linear of a head (:147-149,207-209,221-223): out[n][O] = x[n][HW*C] . w[O][C*HW]^T + bias.  x is
 *                      channels-last (index pixel * C + channel), w keeps nn.Flatten's (channel, pixel) column order -- the kernels
 *                      permute the index, no packed copy of the weights exists.  O <= 16.
 * mz_linear_bwd        g = d loss / d out [n][O]: dx[n][HW*C] (may be NULL), dw[O][C*HW] and db[O] (dw NULL: neither) -- written, or added
 *                      to when accumulate != 0 (the K unroll steps share the heads).  scratch: mz_linear_scratch_bytes() bytes.
 * mz_scale_train_fwd   MuZeroAgent._scale_state (:314-328) per sample over E = H*W*C elements: y = (x - min) / (max - min + 1e-8) as
 *                      float32 and / or 16-bit; stats float32[n][4] = {min, max, argmin, argmax (int bits)} for the backward pass
 * mz_scale_train_bwd   dx = g / r, plus the gradients that flow through min and max to their (first) arg positions, as torch's
 *                      autograd does for tensor.min(dim) / max(dim)
 * mz_planes_conv_fwd   z[n][H*W][cout] += the 3x3 "same" convolution of the A <= 4 action planes (input channels c0 .. c0+A-1 of
 *                      w float32[cout][cin_total][3][3]; the dynamics ConvBlock's torch.cat, :295 / :117-122).  planes float32 with
 *                      ELEMENT strides (sn, sa, sy, sx): the reference passes an expanded (n, A, 1, 1) view.  H*W <= 64.
 * mz_planes_conv_wgrad dw[co][c0+a][tap] (+)= sum_{n,p} dz[n][p][co] * planes[n][a][p + tap]; scratch: mz_planes_wgrad_scratch_bytes()
 */
int mz_cvt16(long long n, const float *src, void *dst, int dtype, void *stream);
/* The two 16-bit packs of a convolution weight w float32[cout][cin_total][ksize][ksize] (its first cin input channels; cout, cin multiples
 * of 64) in ONE launch: fwd (may be NULL) = the tile-contiguous forward operand [tap][cin/64][cout][64] (mz_op.w_layout 1) in fwd_dtype;
 * dgrad (may be NULL) = the same layout of the transposed, tap-flipped filter [tap][cout/64][cin][64] in bf16, whose "same" convolution with
 * dy is the data gradient.  Run before every training iteration's first use of the weight (the optimizer has changed it). */
int mz_pack_conv(int cout, int cin_total, int cin, int ksize, const float *w, void *fwd, int fwd_dtype, void *dgrad, void *stream);
int mz_pool2_train_fwd(int n, int H, int W, int C, const float *x, float *y, void *y16, int dtype, void *stream);
int mz_pool2_train_bwd(int n, int H, int W, int C, const float *dy, float *dx, void *stream);
int mz_linear_fwd(int n, int HW, int C, int O, const float *x, const float *w, const float *bias, float *out, void *stream);
size_t mz_linear_scratch_bytes(int n, int HW, int C, int O);
int mz_linear_bwd(int n, int HW, int C, int O, const float *x, const float *w, const float *g, float *dx, float *dw, float *db, int accumulate,
                  void *scratch, void *stream);
int mz_scale_train_fwd(int n, int E, const float *x, float *y, void *y16, int dtype, float *stats, void *stream);
int mz_scale_train_bwd(int n, int E, const float *x, const float *g, const float *stats, float *dx, void *stream);
int mz_planes_conv_fwd(int n, int H, int W, int A, int cout, int cin_total, int c0, const float *planes, long long sn, long long sa, long long sy,
                       long long sx, const float *w, float *z, void *stream);
size_t mz_planes_wgrad_scratch_bytes(int n, int cout);
int mz_planes_conv_wgrad(int n, int H, int W, int A, int cout, int cin_total, int c0, const float *planes, long long sn, long long sa, long long sy,
                         long long sx, const float *dz, float *dw, int accumulate, void *scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MZB200_H */
