// Batched latent-space MCTS bookkeeping for B200 (sm_100a): flat preallocated per-root node arrays in
// HBM, one warp per tree, the hot upper levels staged in shared memory with coalesced 128-bit loads,
// pUCT argmax over the action lanes by warp shuffle, and backup(sim) + select(sim+1) + latent gather
// (+ root policy extraction on the last simulation) fused into one persistent kernel per simulation.
//
// Replaces (behaviour, not code) src/mcts.py of the reference: _initialize_trees :73-89,
// _expand_root_nodes :91-134, _select_nodes :136-182, _backup :203-234, _compute_results :236-250,
// ucb_action :281-298.  Arithmetic is the reference's: every tensor op a separately rounded fp32 op
// (__fmul_rn/__fadd_rn/__fdiv_rn, no FMA), sqrt / log terms from host-computed double->fp32 tables.
// Tie-break semantics: uniform over exact-equal fp32 maxima in ascending action order, index =
// u32(seed, tree, per-tree pUCT call counter) % count, one draw per pUCT call (mcts.py:294-297 with the
// global mt19937 replaced by the counter-based stream of common.cuh:mz_rng_u32).
#include <math.h>

#include "common.cuh"

namespace {

constexpr int NA = 3;
constexpr int STAGE_NODES = 64;      // nodes of each tree staged in shared memory (slots are allocated in
                                     // expansion order, so low slots are the upper levels of the tree)
constexpr int WARPS_PER_BLOCK = 4;

// Per-tree block in HBM (all 16-byte rows):
//   float4 Q[nodes]   {Q0,Q1,Q2, vsum}
//   float4 P[nodes]   {P0,P1,P2, node value}
//   float4 R[nodes]   {R0,R1,R2, -}
//   int4   NC[nodes]  {N0|child0<<16, N1|child1<<16, N2|child2<<16, fresh}      child 0xFFFF = unexpanded
//   int4   meta[1 + ceil((S+1)/4)]  {nslots, ctr, path_len, leaf: parent | action<<16 } then the path,
//          one int per level: node | action << 16
struct TreeLayout {
    int nodes;        // S + 2
    int path_ints;    // S + 1 rounded up to 4
    __host__ __device__ int meta_off() const { return 4 * nodes; }
    __host__ __device__ int rows() const { return 4 * nodes + 1 + path_ints / 4; }   // 16-byte rows per tree
};
__host__ __device__ inline TreeLayout make_layout(int S)
{
    TreeLayout l;
    l.nodes = S + 2;
    l.path_ints = ((S + 1) + 3) & ~3;
    return l;
}

constexpr int NO_CHILD = 0xFFFF;

struct Params {
    int B, S, sim;               // sim = index of the simulation whose network outputs are being backed up
    TreeLayout lay;
    int4 *trees;                 // [B][rows]
    const float *s_tab, *k_tab;  // [S+1] float32(sqrt(n)), float32(c1 + log((n + c2 + 1) / c2))
    float discount;
    unsigned long long seed;
    const unsigned long long *seed_dev;
    // network outputs of simulation `sim` (backup) -- or of the root prediction (root kernel)
    const float *reward, *value, *pi;   // [B], [B], [B][3]
    const float *noise;                 // [B][3] Dirichlet sample (root kernel only)
    float w_prior, w_noise;             // float32(1 - noise_weight), float32(noise_weight)
    // selection outputs for the next network evaluation
    int *leaf_parent, *leaf_action, *leaf_slot;   // [B]
    const uint4 *latent_store;          // [B][nodes][latent_vec] (16-byte vectors)
    uint4 *dyn_in;                      // [B][latent_vec]: parent latent of the selected leaf
    int latent_vec;
    // results (written when sim == S-1)
    float *out_value;
    long long *out_visits;              // [B][3]
    int *depth_hist;                    // optional [S+1] histogram of selection depths (profiling aid)
};

struct TreeView {
    int4 *g;         // tree block in global memory
    int4 *s;         // staged copy: rows [0,STAGE) of each of the 4 node arrays + meta row + path
    int nodes, stage;
    // node-array row r of array `arr` (0=Q,1=P,2=R,3=NC)
    __device__ __forceinline__ int4 *row(int arr, int slot) const
    {
        return slot < stage ? s + arr * stage + slot : g + arr * nodes + slot;
    }
};

__device__ __forceinline__ float4 ldf(const int4 *p)
{
    int4 v = *p;
    return make_float4(__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z), __int_as_float(v.w));
}
__device__ __forceinline__ void stf(int4 *p, float4 v)
{
    *p = make_int4(__float_as_int(v.x), __float_as_int(v.y), __float_as_int(v.z), __float_as_int(v.w));
}
__device__ __forceinline__ float pick3(float4 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ int pick3i(int4 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
__device__ __forceinline__ void set3(float4 &v, int a, float x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }
__device__ __forceinline__ void set3i(int4 &v, int a, int x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }

// ucb_action(), mcts.py:281-298.  Lane a (< 3) scores action a; max / tie set by warp shuffle + ballot.
__device__ __forceinline__ int puct_select(const Params &p, const TreeView &t, int slot, int tree, int &ctr, int lane)
{
    const float4 q = ldf(t.row(0, slot)), pr = ldf(t.row(1, slot));
    const int4 nc = *t.row(3, slot);
    const int n0 = nc.x & 0xFFFF, n1 = nc.y & 0xFFFF, n2 = nc.z & 0xFFFF;
    const int visit_sum = n0 + n1 + n2;                               // :285 (this node's children)
    const float s = p.s_tab[visit_sum], k = p.k_tab[visit_sum];       // :286-289 python doubles -> fp32
    const int a = lane < NA ? lane : NA - 1;
    const int n = a == 0 ? n0 : (a == 1 ? n1 : n2);
    float u = __fmul_rn(pick3(pr, a), s);
    u = __fdiv_rn(u, (float)(1 + n));
    u = __fmul_rn(u, k);
    float score = __fadd_rn(pick3(q, a), u);
    if (lane >= NA) score = -INFINITY;
    float best = fmaxf(score, __shfl_xor_sync(0xffffffffu, score, 1));
    best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
    best = __shfl_sync(0xffffffffu, best, 0);                         // lanes 0-3 hold the max; broadcast
    const unsigned cand = __ballot_sync(0xffffffffu, lane < NA && score == best);   // :294-296
    const unsigned u32 = mzb::mz_rng_u32(p.seed, (uint32_t)tree, (uint32_t)ctr);    // :297, one draw per call
    ctr += 1;
    if (cand == 0u) return 0;                                         // every score NaN (non-finite network output): nothing compares equal; action 0
    int pick = (int)(u32 % (unsigned)__popc(cand));
    return __fns(cand, 0, pick + 1);                                  // pick-th set bit, ascending action order
}

__device__ __forceinline__ void node_reset(const TreeView &t, int slot, int fresh)
{
    stf(t.row(0, slot), make_float4(0.f, 0.f, 0.f, 0.f));
    stf(t.row(1, slot), make_float4(0.f, 0.f, 0.f, 0.f));
    stf(t.row(2, slot), make_float4(0.f, 0.f, 0.f, 0.f));
    *t.row(3, slot) = make_int4(NO_CHILD << 16, NO_CHILD << 16, NO_CHILD << 16, fresh);
}

// _select_nodes(), mcts.py:136-182 (from the root), also used for the first action at the root (:124).
// Executed by lane 0's view of the data but with the whole warp converged (puct_select uses shuffles).
__device__ __forceinline__ void select_leaf(const Params &p, const TreeView &t, int4 *meta, int *path, int tree, int lane,
                                            int &nslots, int &ctr, int &parent, int &action, int &leaf, int &len,
                                            bool root_first_action)
{
    int cur = 0;
    len = 0;
    for (;;) {
        const int a = puct_select(p, t, cur, tree, ctr, lane);
        int4 nc = *t.row(3, cur);
        int c = (pick3i(nc, a) >> 16) & 0xFFFF;
        const bool expanded = c != NO_CHILD && t.row(3, c)->w == 0;   // subtree["expanded"] :163
        if (expanded && !root_first_action) {
            if (lane == 0) path[len] = cur | (a << 16);               // :165 (edge reward is re-read at backup)
            ++len;
            cur = c;
            continue;
        }
        if (c == NO_CHILD) {                                          // placeholder -> new slot :167-175
            c = nslots < p.lay.nodes ? nslots++ : p.lay.nodes - 1;    // a search allocates at most S+1 slots; stay in bounds if misused
            if (lane == 0) { set3i(nc, a, (pick3i(nc, a) & 0xFFFF) | (c << 16)); *t.row(3, cur) = nc; }
        }
        __syncwarp();
        if (lane == 0) node_reset(t, c, root_first_action ? 1 : 0);   // sim-0 leaf keeps "expanded": False (:121)
        __syncwarp();
        parent = cur; action = a; leaf = c;
        break;
    }
    (void)meta;
}

// _backup(), mcts.py:203-234
__device__ __forceinline__ void backup(const Params &p, const TreeView &t, const int *path, int tree, int lane,
                                       int parent, int action, int leaf, int len)
{
    if (lane != 0) return;
    const float r = p.reward[tree], v0 = p.value[tree];
    float4 pr = make_float4(p.pi[tree * 3 + 0], p.pi[tree * 3 + 1], p.pi[tree * 3 + 2], v0);
    float4 rr = ldf(t.row(2, parent));
    set3(rr, action, r);                                              // :215
    stf(t.row(2, parent), rr);
    stf(t.row(0, leaf), make_float4(0.f, 0.f, 0.f, v0));              // :216, :219-225
    stf(t.row(1, leaf), pr);
    stf(t.row(2, leaf), make_float4(0.f, 0.f, 0.f, 0.f));
    int4 lnc = *t.row(3, leaf);
    *t.row(3, leaf) = make_int4(NO_CHILD << 16, NO_CHILD << 16, NO_CHILD << 16, lnc.w);
    float v = v0;
    for (int k = len; k >= 0; --k) {                                  // :227-234, leaf edge first
        const int node = k == len ? parent : (path[k] & 0xFFFF);
        const int a = k == len ? action : (path[k] >> 16);
        const float re = pick3(ldf(t.row(2, node)), a);
        v = __fadd_rn(__fmul_rn(v, p.discount), re);
        float4 q = ldf(t.row(0, node));
        int4 nc = *t.row(3, node);
        const int e = pick3i(nc, a);
        const int n = e & 0xFFFF;
        q.w = __fadd_rn(q.w, v);                                      // node["value"] += value  :232
        float qa = __fadd_rn(__fmul_rn((float)n, pick3(q, a)), v);    // (N*Q + value)/(N+1)     :233
        set3(q, a, __fdiv_rn(qa, (float)(n + 1)));
        set3i(nc, a, (e & ~0xFFFF) | (n + 1));                        // :234
        stf(t.row(0, node), q);
        *t.row(3, node) = nc;
    }
}

template <bool kRoot>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) tree_kernel(Params p)
{
    if (p.seed_dev) p.seed = *p.seed_dev;
    extern __shared__ int4 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stage = min(STAGE_NODES, p.lay.nodes);
    const int meta_rows = 1 + p.lay.path_ints / 4;
    const int stage_rows = 4 * stage + meta_rows;
    int4 *s = smem + warp * stage_rows;
    const int warps_total = gridDim.x * WARPS_PER_BLOCK;

    for (int tree = blockIdx.x * WARPS_PER_BLOCK + warp; tree < p.B; tree += warps_total) {
        int4 *g = p.trees + (size_t)tree * p.lay.rows();
        TreeView t{g, s, p.lay.nodes, stage};
        int4 *meta = s + 4 * stage;
        int *path = reinterpret_cast<int *>(meta + 1);
        int nslots, ctr, len, parent, action, leaf;

        if (kRoot) {
            // _initialize_trees + _expand_root_nodes: :73-134
            nslots = 1; ctr = 0;
            if (lane == 0) {
                node_reset(t, 0, 0);
                const float v_root = p.value[tree];
                float4 pr;
                pr.x = __fadd_rn(__fmul_rn(p.w_prior, p.pi[tree * 3 + 0]), __fmul_rn(p.w_noise, p.noise[tree * 3 + 0]));   // :119
                pr.y = __fadd_rn(__fmul_rn(p.w_prior, p.pi[tree * 3 + 1]), __fmul_rn(p.w_noise, p.noise[tree * 3 + 1]));
                pr.z = __fadd_rn(__fmul_rn(p.w_prior, p.pi[tree * 3 + 2]), __fmul_rn(p.w_noise, p.noise[tree * 3 + 2]));
                pr.w = v_root;
                stf(t.row(1, 0), pr);
                stf(t.row(0, 0), make_float4(0.f, 0.f, 0.f, v_root));     // root value :110
            }
            __syncwarp();
            select_leaf(p, t, meta, path, tree, lane, nslots, ctr, parent, action, leaf, len, true);   // :124
        } else {
            // stage the upper levels + meta/path: coalesced 16-byte loads
            for (int arr = 0; arr < 4; ++arr)
                for (int i = lane; i < stage; i += 32) s[arr * stage + i] = g[arr * p.lay.nodes + i];
            for (int i = lane; i < meta_rows; i += 32) meta[i] = g[p.lay.meta_off() + i];
            __syncwarp();
            const int4 m = meta[0];
            nslots = m.x; ctr = m.y; len = m.z; parent = m.w & 0xFFFF; action = (m.w >> 16) & 0x3; leaf = (int)((unsigned)m.w >> 18);
            backup(p, t, path, tree, lane, parent, action, leaf, len);
            __syncwarp();
            if (p.sim + 1 < p.S) {
                select_leaf(p, t, meta, path, tree, lane, nslots, ctr, parent, action, leaf, len, false);
            } else if (lane == 0) {
                // _compute_results :236-250: value = float32(double(root value) / num_simulations)
                const float4 q = ldf(t.row(0, 0));
                const int4 nc = *t.row(3, 0);
                p.out_value[tree] = (float)((double)q.w / (double)p.S);
                p.out_visits[tree * 3 + 0] = nc.x & 0xFFFF;
                p.out_visits[tree * 3 + 1] = nc.y & 0xFFFF;
                p.out_visits[tree * 3 + 2] = nc.z & 0xFFFF;
            }
        }
        const bool selecting = kRoot || p.sim + 1 < p.S;
        if (lane == 0) {
            meta[0] = make_int4(nslots, ctr, len, parent | (action << 16) | (leaf << 18));
            if (selecting) {
                p.leaf_parent[tree] = parent; p.leaf_action[tree] = action; p.leaf_slot[tree] = leaf;
                if (p.depth_hist) atomicAdd(p.depth_hist + len, 1);
            }
        }
        __syncwarp();
        // write the staged rows back (coalesced)
        for (int arr = 0; arr < 4; ++arr)
            for (int i = lane; i < stage; i += 32) g[arr * p.lay.nodes + i] = s[arr * stage + i];
        for (int i = lane; i < meta_rows; i += 32) g[p.lay.meta_off() + i] = meta[i];
        // gather the selected leaf's parent latent into the dynamics-network input (128-bit copies)
        if (selecting && p.dyn_in) {
            const uint4 *src = p.latent_store + ((size_t)tree * p.lay.nodes + parent) * p.latent_vec;
            uint4 *dst = p.dyn_in + (size_t)tree * p.latent_vec;
            // batches of 8 loads in flight per lane before the stores (src and dst may alias as far as the compiler knows, so a plain
            // copy loop is one dependent L2 round trip per 512 bytes: 20 of them for a 10 KB latent, most of a small batch's tree step)
            for (int i0 = lane; i0 < p.latent_vec; i0 += 32 * 8) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (i0 + 32 * k < p.latent_vec) v[k] = __ldcs(src + i0 + 32 * k);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (i0 + 32 * k < p.latent_vec) dst[i0 + 32 * k] = v[k];
            }
        }
        __syncwarp();
    }
}

int launch(Params &p, bool root, cudaStream_t st)
{
    const int stage = p.lay.nodes < STAGE_NODES ? p.lay.nodes : STAGE_NODES;
    const int stage_rows = 4 * stage + 1 + p.lay.path_ints / 4;
    const size_t smem = (size_t)WARPS_PER_BLOCK * stage_rows * sizeof(int4);
    int blocks = (p.B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int max_blocks = mzb::kNumSMs * 8;                  // persistent: at most 8 CTAs per SM, loop over trees
    if (blocks > max_blocks) blocks = max_blocks;
    if (smem > 48 * 1024) {
        MZB_CUDA(cudaFuncSetAttribute(tree_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MZB_CUDA(cudaFuncSetAttribute(tree_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (root) tree_kernel<true><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(p);
    else tree_kernel<false><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(p);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

size_t mz_tree_bytes(int num_simulations)
{
    return num_simulations > 0 ? (size_t)make_layout(num_simulations).rows() * sizeof(int4) : 0;
}

int mz_tree_nodes(int num_simulations) { return num_simulations + 2; }

int mz_puct_tables(int num_simulations, double c1, double c2, float *s_tab_host, float *k_tab_host)
{
    MZB_CHECK_ARG(num_simulations > 0 && s_tab_host && k_tab_host, "bad argument");
    for (int n = 0; n <= num_simulations; ++n) {
        s_tab_host[n] = (float)sqrt((double)n);                               // math.sqrt(visit_sum)
        k_tab_host[n] = (float)(c1 + log(((double)n + c2 + 1.0) / c2));       // c1 + math.log((visit_sum + c2 + 1) / c2)
    }
    return 0;
}

static int fill(Params &p, const mz_tree_args *a)
{
    MZB_CHECK_ARG(a && a->B > 0 && a->num_simulations > 0 && a->num_simulations <= 8190, "bad B / num_simulations (at most 8190: node slots are 13-bit fields of the selection record)");
    MZB_CHECK_ARG(a->trees && a->s_tab && a->k_tab && a->value && a->pi, "null pointer");
    MZB_CHECK_ARG(a->leaf_parent && a->leaf_action && a->leaf_slot, "null selection outputs");
    MZB_CHECK_ARG((a->latent_bytes % 16) == 0, "latent_bytes must be a multiple of 16");
    MZB_CHECK_ARG(!a->dyn_in || a->latent_store, "dyn_in needs latent_store");
    p.B = a->B; p.S = a->num_simulations; p.sim = a->sim;
    p.lay = make_layout(a->num_simulations);
    p.trees = (int4 *)a->trees;
    p.s_tab = a->s_tab; p.k_tab = a->k_tab;
    p.discount = (float)a->discount;
    p.seed = a->seed;
    p.seed_dev = (const unsigned long long *)a->seed_dev;
    p.reward = a->reward; p.value = a->value; p.pi = a->pi; p.noise = a->noise;
    p.w_prior = (float)(1.0 - a->noise_weight); p.w_noise = (float)a->noise_weight;
    p.leaf_parent = a->leaf_parent; p.leaf_action = a->leaf_action; p.leaf_slot = a->leaf_slot;
    p.latent_store = (const uint4 *)a->latent_store; p.dyn_in = (uint4 *)a->dyn_in; p.latent_vec = (int)(a->latent_bytes / 16);
    p.out_value = a->out_value; p.out_visits = (long long *)a->out_visits;
    p.depth_hist = a->depth_hist;
    return 0;
}

int mz_tree_root(const mz_tree_args *a, void *stream)
{
    Params p;
    if (int rc = fill(p, a)) return rc;
    MZB_CHECK_ARG(a->noise, "root needs the Dirichlet noise");
    return launch(p, true, (cudaStream_t)stream);
}

int mz_tree_step(const mz_tree_args *a, void *stream)
{
    Params p;
    if (int rc = fill(p, a)) return rc;
    MZB_CHECK_ARG(a->reward, "step needs the leaf rewards");
    MZB_CHECK_ARG(a->sim >= 0 && a->sim < a->num_simulations, "sim out of range");
    MZB_CHECK_ARG(a->sim + 1 < a->num_simulations || (a->out_value && a->out_visits), "last simulation needs the result buffers");
    return launch(p, false, (cudaStream_t)stream);
}

}  // extern "C"
