/*
 * oracle/mcts_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, flat slot arrays) of the tree bookkeeping of the reference's
 * latent-space MCTS:
 *     /root/reference/src/mcts.py
 *         MCTSSearchVec._initialize_trees   :73-89
 *         MCTSSearchVec._expand_root_nodes  :91-134   (root priors + Dirichlet noise, first action)
 *         MCTSSearchVec._select_nodes       :136-182  (pUCT descent, leaf creation)
 *         MCTSSearchVec._backup             :203-234  (edge creation, discounted backup, Q/N update)
 *         MCTSSearchVec._compute_results    :236-250
 *         MCTSSearchVec.ucb_action          :281-298
 * The network calls (_expand_nodes :184-201) are NOT here: the driver (tests / bench) evaluates
 * the networks and feeds (reward, value, policy-probabilities) back in lock step.
 *
 * Randomness: the reference draws Dirichlet noise (:114) and the tie-break index (:297) from the
 * global CPU mt19937, in a data-dependent serial order that a GPU cannot reproduce.  Parity is
 * therefore defined through injection: the noise is an explicit (B,3)
 * input, and the tie-break draw is u32(seed, tree, per-tree call counter) % count, one draw per
 * pUCT call whether or not there is a tie.  tests/golden/gen_golden.py patches exactly those two
 * draws in the UNMODIFIED reference and records its results; tests/test_oracle_mcts.py pins this
 * file against those records (visit counts identical, root value bit-identical).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NA 3 /* actions = [0,1,2] (config.yaml:6); edge key 3 of the reference is never used */

typedef struct {
    int32_t N[NA];
    float Q[NA], P[NA], R[NA];
    int32_t child[NA]; /* slot of the child node, -1 = unexpanded placeholder */
    float vsum;        /* node["value"], accumulates every backed-up return (:232) */
    int32_t fresh;     /* the sim-0 leaf keeps "expanded": False (:121) -> re-expanded on 2nd visit */
} Node;

typedef struct {
    int B, S, cap;
    double c1, c2;
    float discount;
    uint64_t seed;
    int sim;
    Node *nodes;       /* [B][cap] */
    int32_t *nslots;   /* [B] */
    uint32_t *ctr;     /* [B] pUCT call counter (RNG stream position) */
    int32_t *path_n;   /* [B][S+1] nodes on the current trajectory (root first) */
    int32_t *path_a;   /* [B][S+1] */
    float *path_r;     /* [B][S+1] edge reward captured at selection time (:165) */
    int32_t *path_len; /* [B] number of (node,action) pairs above the leaf's parent edge */
    int32_t *leaf_parent, *leaf_action, *leaf_slot; /* [B] */
} MTO;

/* counter-based tie-break stream shared with the CUDA tree kernel (csrc/tree.cu: mz_rng_u32) */
uint32_t mto_rng_u32(uint64_t seed, uint32_t tree, uint32_t ctr)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)tree) << 32) | (uint64_t)ctr);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

MTO *mto_create(int B, int S, double c1, double c2, double discount, uint64_t seed)
{
    MTO *t = (MTO *)calloc(1, sizeof(MTO));
    t->B = B; t->S = S; t->cap = S + 2;
    t->c1 = c1; t->c2 = c2; t->discount = (float)discount; t->seed = seed;
    t->nodes = (Node *)calloc((size_t)B * t->cap, sizeof(Node));
    t->nslots = (int32_t *)calloc(B, sizeof(int32_t));
    t->ctr = (uint32_t *)calloc(B, sizeof(uint32_t));
    t->path_n = (int32_t *)calloc((size_t)B * (S + 1), sizeof(int32_t));
    t->path_a = (int32_t *)calloc((size_t)B * (S + 1), sizeof(int32_t));
    t->path_r = (float *)calloc((size_t)B * (S + 1), sizeof(float));
    t->path_len = (int32_t *)calloc(B, sizeof(int32_t));
    t->leaf_parent = (int32_t *)calloc(B, sizeof(int32_t));
    t->leaf_action = (int32_t *)calloc(B, sizeof(int32_t));
    t->leaf_slot = (int32_t *)calloc(B, sizeof(int32_t));
    return t;
}

void mto_destroy(MTO *t)
{
    if (!t) return;
    free(t->nodes); free(t->nslots); free(t->ctr); free(t->path_n); free(t->path_a);
    free(t->path_r); free(t->path_len); free(t->leaf_parent); free(t->leaf_action);
    free(t->leaf_slot); free(t);
}

static void node_init(Node *n)
{
    for (int a = 0; a < NA; ++a) { n->N[a] = 0; n->Q[a] = 0.0f; n->P[a] = 0.0f; n->R[a] = 0.0f; n->child[a] = -1; }
    n->vsum = 0.0f;
    n->fresh = 0;
}

/* ucb_action(): mcts.py:281-298.  sqrt and the log term are Python doubles that are rounded to
 * float32 when they meet the float32 tensor; every tensor op is a separately rounded fp32 op. */
static int ucb_action(MTO *t, int b, const Node *n)
{
    int visit_sum = n->N[0] + n->N[1] + n->N[2];                     /* :285 */
    double log_term = ((double)visit_sum + t->c2 + 1.0) / t->c2;     /* :286 */
    float s = (float)sqrt((double)visit_sum);
    float k = (float)(t->c1 + log(log_term));
    volatile float score[NA];
    for (int a = 0; a < NA; ++a) {                                   /* :288-289 */
        volatile float u = n->P[a] * s;
        u = u / (float)(1 + n->N[a]);
        u = u * k;
        score[a] = n->Q[a] + u;
    }
    float best = score[0];
    for (int a = 1; a < NA; ++a) if (score[a] > best) best = score[a];
    int cand[NA], nc = 0;
    for (int a = 0; a < NA; ++a) if (score[a] == best) cand[nc++] = a; /* :294-296 */
    uint32_t u = mto_rng_u32(t->seed, (uint32_t)b, t->ctr[b]++);     /* :297, one draw per call */
    return cand[u % (uint32_t)nc];
}

/* _expand_root_nodes(): mcts.py:91-134.  v_root = inverted_softmax_expectation(value logits),
 * pi = softmax(policy logits), noise = Dirichlet sample, all float32, supplied by the driver. */
void mto_root(MTO *t, const float *v_root, const float *pi, const float *noise, double noise_weight,
              int32_t *parent_slot, int32_t *action, int32_t *leaf_slot)
{
    const float w1 = (float)(1.0 - noise_weight), w = (float)noise_weight;
    t->sim = 0;
    for (int b = 0; b < t->B; ++b) {
        Node *root = t->nodes + (size_t)b * t->cap;
        node_init(root);
        t->nslots[b] = 1;
        t->ctr[b] = 0;
        root->vsum = v_root[b];                                      /* :110 */
        for (int a = 0; a < NA; ++a) {                               /* :119 */
            volatile float x = w1 * pi[b * NA + a];
            volatile float y = w * noise[b * NA + a];
            root->P[a] = x + y;
        }
        int a0 = ucb_action(t, b, root);                             /* :124 */
        int c = t->nslots[b]++;
        Node *leaf = root + c;
        node_init(leaf);
        leaf->fresh = 1;                                             /* placeholder stays "expanded": False */
        root->child[a0] = c;
        t->path_len[b] = 0;                                          /* trajectories.append([]) :133 */
        t->leaf_parent[b] = 0; t->leaf_action[b] = a0; t->leaf_slot[b] = c;
        parent_slot[b] = 0; action[b] = a0; leaf_slot[b] = c;
    }
}

/* _select_nodes(): mcts.py:136-182 */
void mto_select(MTO *t, int32_t *parent_slot, int32_t *action, int32_t *leaf_slot)
{
    for (int b = 0; b < t->B; ++b) {
        Node *base = t->nodes + (size_t)b * t->cap;
        int32_t *pn = t->path_n + (size_t)b * (t->S + 1), *pa = t->path_a + (size_t)b * (t->S + 1);
        float *pr = t->path_r + (size_t)b * (t->S + 1);
        int cur = 0, len = 0;
        for (;;) {
            Node *n = base + cur;
            int a = ucb_action(t, b, n);
            int c = n->child[a];
            if (c >= 0 && !base[c].fresh) {                          /* subtree["expanded"] :163 */
                pn[len] = cur; pa[len] = a; pr[len] = n->R[a]; ++len;  /* :165 */
                cur = c;
                continue;
            }
            if (c < 0) {                                             /* placeholder -> new node :167-175 */
                c = t->nslots[b]++;
                n->child[a] = c;
            }
            node_init(base + c);                                     /* fresh re-expansion wipes the old dict */
            t->leaf_parent[b] = cur; t->leaf_action[b] = a; t->leaf_slot[b] = c;
            break;
        }
        t->path_len[b] = len;
        parent_slot[b] = t->leaf_parent[b]; action[b] = t->leaf_action[b]; leaf_slot[b] = t->leaf_slot[b];
    }
}

/* _backup(): mcts.py:203-234.  reward/value = inverted_softmax_expectation of the dynamics /
 * prediction logits of the leaf, pi = softmax(policy logits of the leaf). */
void mto_backup(MTO *t, const float *reward, const float *value, const float *pi)
{
    for (int b = 0; b < t->B; ++b) {
        Node *base = t->nodes + (size_t)b * t->cap;
        int32_t *pn = t->path_n + (size_t)b * (t->S + 1), *pa = t->path_a + (size_t)b * (t->S + 1);
        float *pr = t->path_r + (size_t)b * (t->S + 1);
        int parent = t->leaf_parent[b], a = t->leaf_action[b], c = t->leaf_slot[b];
        Node *leaf = base + c;
        base[parent].R[a] = reward[b];                               /* :215 */
        leaf->vsum = value[b];                                       /* :216 */
        for (int i = 0; i < NA; ++i) {                               /* :219-225 */
            leaf->N[i] = 0; leaf->Q[i] = 0.0f; leaf->P[i] = pi[b * NA + i]; leaf->R[i] = 0.0f;
            leaf->child[i] = -1;
        }
        int len = t->path_len[b];
        pn[len] = parent; pa[len] = a; pr[len] = reward[b]; ++len;   /* :227 */
        volatile float v = value[b];
        for (int k = len - 1; k >= 0; --k) {                         /* :230-234 */
            Node *n = base + pn[k];
            int e = pa[k];
            v = v * t->discount;
            v = v + pr[k];
            n->vsum = n->vsum + v;
            volatile float q = (float)n->N[e] * n->Q[e];
            q = q + v;
            n->Q[e] = q / (float)(n->N[e] + 1);
            n->N[e] += 1;
        }
    }
    t->sim += 1;
}

/* _compute_results(): mcts.py:236-250.  value = float32(double(root value) / num_simulations). */
void mto_results(MTO *t, float *value, int64_t *visits)
{
    for (int b = 0; b < t->B; ++b) {
        Node *root = t->nodes + (size_t)b * t->cap;
        for (int a = 0; a < NA; ++a) visits[b * NA + a] = root->N[a];
        value[b] = (float)((double)root->vsum / (double)t->S);
    }
}

/* introspection for the lock-step tests */
int mto_path(MTO *t, int b, int32_t *nodes, int32_t *acts)
{
    int len = t->path_len[b];
    for (int k = 0; k < len; ++k) { nodes[k] = t->path_n[(size_t)b * (t->S + 1) + k]; acts[k] = t->path_a[(size_t)b * (t->S + 1) + k]; }
    return len;
}
int mto_nslots(MTO *t, int b) { return t->nslots[b]; }
void mto_node(MTO *t, int b, int slot, int32_t *N, float *Q, float *P, float *R, int32_t *child, float *vsum)
{
    Node *n = t->nodes + (size_t)b * t->cap + slot;
    for (int a = 0; a < NA; ++a) { N[a] = n->N[a]; Q[a] = n->Q[a]; P[a] = n->P[a]; R[a] = n->R[a]; child[a] = n->child[a]; }
    *vsum = n->vsum;
}
