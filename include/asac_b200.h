/*
 * asac_b200.h — C ABI of libasac_b200.so (sm_100a).
 *
 * The reference (BlueFisher/Advanced-Soft-Actor-Critic) is 100 % Python and has no FFI;
 * every entry point below replaces a Python function of the reference's hot path and is
 * what a ctypes binding inside the reference would call.  File:line
 * citations are relative to the reference checkout.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (a torch tensor's data_ptr())
 *     unless the parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises
 *     unless stated;
 *   - the return value is 0 on success, otherwise a negative ASAC_E* code or a positive
 *     cudaError_t; asac_last_error() returns a static description;
 *   - no entry point falls back to the CPU.
 */
#ifndef ASAC_B200_H
#define ASAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASAC_OK 0
#define ASAC_EINVAL (-1)       /* bad argument (shape, alignment, unsupported size) */
#define ASAC_EUNSUPPORTED (-2) /* configuration outside the fused path              */

#define ASAC_MAX_COLUMNS 16
#define ASAC_MAX_BRANCHES 8   /* discrete action branches */
#define ASAC_MAX_NSTEP 16
#define ASAC_MAX_DEPTH 4
#define ASAC_MAX_ENSEMBLE 8
#define ASAC_MAX_PEERS 16

const char *asac_last_error(void);
int asac_version(void);
/* number of kernels this library has launched since load / since the last reset
 * (bench.py's gpu_launches) */
int64_t asac_launch_count(void);
void asac_reset_launch_count(void);
/* Programmatic dependent launch of the step's kernels (default on; ASAC_PDL=0 in the environment turns it off):
 * the critic backward, the policy backward and the post pass start their predecessor-independent part while the
 * kernel ahead of them drains and wait (griddepcontrol.wait) in front of the first dependent read.  Takes effect
 * at the next launch / graph capture; returns the previous setting.  Results do not depend on it. */
int asac_set_pdl(int on);

/* ------------------------------------------------------------------------------------
 * Sum tree  — replaces SumTree (algorithm/replay_buffer.py:145-242).
 *
 * HBM layout: float32[2*capacity], 1-based heap: node 1 is the root, the children of
 * node i are 2i and 2i+1 (one aligned float2), leaf of data slot j is node capacity+j.
 * nodes[1:] is element-for-element the reference's `_tree` array (root at 0, children
 * 2i+1/2i+2), so `<ckpt>-rb_tree.npy` files interchange by an offset of one.
 * Every parent holds fp32 (left + right), exactly as replay_buffer.py:181.
 * ---------------------------------------------------------------------------------- */

/* SumTree.update (replay_buffer.py:172-183): nodes[capacity+slot[i]] = p[i] (duplicates:
 * the last one wins, like NumPy fancy assignment), then every ancestor is recomputed.
 * k may be any size; it is processed in sequential chunks of <= 1024. */
int asac_tree_update(float *nodes, int64_t capacity, const int64_t *slots, const float *p,
                     int64_t k, void *stream);

/* Full bottom-up recomputation of the internal nodes from the leaves (used after
 * SumTree.load, replay_buffer.py:225-227). */
int asac_tree_rebuild(float *nodes, int64_t capacity, void *stream);

/* SumTree.max (replay_buffer.py:237-239): max over the leaves -> out[0]. */
int asac_tree_leaf_max(const float *nodes, int64_t capacity, float *out, void *stream);

/* SumTree.sample (replay_buffer.py:185-205) with the uniforms injected:
 * seg = total/B in fp32; v_i = i*seg + ((i+1)*seg - i*seg) * u[i] in float64; descent
 * "left iff v <= left || right == 0" comparing in float64.  When `unit_uniform` is NULL the
 * uniforms come from Philox4x32-10 keyed by (seed, draw_counter[0]).
 * out_slot[i] = leaf - capacity (int32), out_p[i] = the leaf's priority. */
int asac_tree_sample(const float *nodes, int64_t capacity, int batch, const double *unit_uniform,
                     uint64_t seed, const int64_t *draw_counter, int32_t *out_slot, float *out_p,
                     void *stream);

/* ------------------------------------------------------------------------------------
 * Prioritized replay — replaces PrioritizedReplayBuffer (replay_buffer.py:245-477).
 * `store_ids` is DataStorage's `_id` column, int64[capacity] (replay_buffer.py:36).
 * `per_state` is double[4] on the device: {beta, beta_increment, min sampling probability of the last
 * batch, nan_flag}.
 * ---------------------------------------------------------------------------------- */

/* One iteration of _prefetch_loop's sampling block (replay_buffer.py:347-354): tree
 * sample, data ids, beta += increment (capped at 1), IS weights (w/min w)^-beta computed in
 * float64 and rounded to fp32.  batch <= 1024.  draw_counter[0] is advanced by one when the
 * uniforms are generated on the device. */
int asac_per_sample(const float *nodes, int64_t capacity, const int64_t *store_ids, int batch,
                    const double *unit_uniform, uint64_t seed, int64_t *draw_counter,
                    double *per_state, int32_t *out_slot, int64_t *out_data_id, float *out_p,
                    float *out_is_weight, void *stream);

/* Sharded replay (one tree per GPU; new capability, the reference has one buffer): recomputes the batch's IS
 * weights against the smallest sampling probability over ALL shards' batches, global_min[0] = MIN over ranks of
 * per_state[2] after asac_per_sample — replay_buffer.py:352-354 applied to the union of the shards' draws:
 * w = ((p / total_of_this_shard) / global_min)^-beta.  With one shard it reproduces asac_per_sample's weights. */
int asac_per_shard_weights(const float *nodes, int batch, const float *p, const double *per_state,
                           const double *global_min, float *out_is_weight, void *stream);

/* PrioritizedReplayBuffer.update (replay_buffer.py:412-427): p = clip(|td|, min, max)^alpha in
 * fp32, skipped where store_ids[id % capacity] != id, then SumTree.update.  A NaN td sets
 * per_state[3] = 1 and leaves the tree untouched (the reference raises).  k <= 1024.
 * When `precomputed_p` is non-zero `td` already holds priorities (add() path). */
int asac_per_update(float *nodes, int64_t capacity, const int64_t *store_ids,
                    const int64_t *data_ids, const float *td, int k, float td_min, float td_max,
                    float alpha, int precomputed_p, double *per_state, void *stream);

/* PrioritizedReplayBuffer.add's bookkeeping (replay_buffer.py:293-307 + DataStorage.add
 * :43-54): writes store_ids for T new rows starting at `first_id`, gives them priority
 * max_p (device scalar; td_error_max when the buffer was empty), zero for the last
 * `ignore_size` rows and for rows landing in the last `ignore_size` ring slots, and updates
 * the tree.  The row payload itself is written by asac_storage_write_rows. */
int asac_per_add(float *nodes, int64_t capacity, int64_t *store_ids, int64_t first_id, int64_t T,
                 const float *max_p, int ignore_size, void *stream);

/* ------------------------------------------------------------------------------------
 * Transition storage — replaces DataStorage (replay_buffer.py:21-142) and the learner-side
 * padding of SAC_Base._sample_from_replay_buffer (sac_base.py:2435-2453).
 * ---------------------------------------------------------------------------------- */

/* DataStorage.add payload (replay_buffer.py:45-50): ring[(first_id+i) % (10*capacity) %
 * capacity] = rows[i] for one column of `row_bytes` bytes per row. */
int asac_storage_write_rows(void *ring, int64_t capacity, int64_t first_id, const void *rows,
                            int64_t T, int64_t row_bytes, void *stream);

/* DataStorage.add for every key of one episode in a single launch: the caller stages the
 * episode's columns one after the other in ONE device buffer (one H2D copy from pinned host
 * memory); col[c].rows points at column c's [T, row_bytes] block inside it. */
typedef struct {
    int32_t n_columns;
    struct {
        void *ring;        /* [capacity, row_bytes] */
        const void *rows;  /* [T, row_bytes]        */
        int64_t row_bytes;
    } col[ASAC_MAX_COLUMNS];
} AsacWriteTable;

int asac_storage_write_table(const AsacWriteTable *table_host, int64_t capacity, int64_t first_id,
                             int64_t T, void *stream);

/* PrioritizedReplayBuffer.add for a HOST-resident episode as one call (replay_buffer.py:293-315 +
 * DataStorage.add :30-56): packs the columns into a pinned staging buffer owned by the handle,
 * one cudaMemcpyAsync, then asac_tree_leaf_max (unless the buffer was empty: td_error_max is used,
 * :296-299), asac_storage_write_table and asac_per_add on `stream` — for episodes of <= 1024 rows and <= 256 KB
 * those three steps are ONE kernel (k_ingest_small: every CTA scans its share of the leaves and takes a ticket,
 * the CTA with the last ticket copies the rows and inserts them; bit-identical tree and rings).  The handle owns only its
 * staging buffers/events; `rings` (one [capacity, row_bytes] device array per key, in the order
 * of `host_columns`), `nodes`, `store_ids`, `max_p_scratch` (device float) and `td_max_dev`
 * (device float holding td_error_max) stay the caller's and must outlive the handle.
 * host_columns[c] is [T, row_bytes[c]] contiguous host memory; T <= capacity; first_id is
 * DataStorage._id before the call. */
typedef struct AsacIngest AsacIngest;
int asac_ingest_create(AsacIngest **out, int64_t capacity, int n_columns, void *const *rings,
                       const int64_t *row_bytes, float *nodes, int64_t *store_ids, float *max_p_scratch,
                       const float *td_max_dev);
int asac_ingest_add(AsacIngest *h, const void *const *host_columns, int64_t T, int64_t first_id,
                    int ignore_size, int buffer_empty, void *stream);
int64_t asac_ingest_row_bytes(const AsacIngest *h);
void asac_ingest_destroy(AsacIngest *h);

enum {
    ASAC_ROLE_COPY = 0,      /* obs, last_mask: copied as stored                           */
    ASAC_ROLE_INDEX = 1,     /* int32 episode index: -1 on padded rows (sac_base.py:2445)  */
    ASAC_ROLE_ACTION = 2,    /* float32[A]: padding_action on padded rows (:2449)          */
    ASAC_ROLE_REWARD = 3,    /* float32: 0 on padded rows (:2450)                          */
    ASAC_ROLE_DONE = 4,      /* bool: true on padded rows (:2451)                          */
    ASAC_ROLE_MU_PROB = 5,   /* float32[A]: 1 on padded rows (:2452)                       */
    ASAC_ROLE_HIDDEN = 6     /* float32[*]: 0 on padded rows (:2453)                       */
};

typedef struct {
    const void *ring;      /* [capacity, row_bytes] */
    void *out;             /* destination base                                              */
    int32_t row_bytes;     /* bytes per stored row                                          */
    int32_t out_stride;    /* bytes between consecutive (b, t) rows in `out`                */
    int32_t out_offset;    /* byte offset inside the destination row (state concatenation)  */
    int32_t role;          /* ASAC_ROLE_*                                                    */
} AsacColumn;

typedef struct {
    int32_t n_columns;
    int32_t index_column;  /* which column holds the int32 episode index */
    AsacColumn col[ASAC_MAX_COLUMNS];
} AsacColumnTable;

/* get_storage_data over the windows data_id + [-prev_n, post_n] (replay_buffer.py:356-362)
 * fused with the padding rule: a row is padding when stored_index(b,t) -
 * stored_index(b,prev_n) != t - prev_n.  out_padding_mask is uint8[batch, L].
 * padding_action is float32[A] (sac_base.py:291-294). */
int asac_storage_gather(const AsacColumnTable *table_host, int64_t capacity,
                        const int64_t *data_ids, int batch, int prev_n, int post_n,
                        const float *padding_action, uint8_t *out_padding_mask, void *stream);

/* update_transitions (replay_buffer.py:429-434) for the write-backs of sac_base.py:2586-2605:
 * ring[(data_id[b] + first_offset + t) % capacity] = rows[b, t] for t < n_rows, skipped on
 * padded rows and where store_ids no longer holds that id. */
int asac_storage_scatter(void *ring, int64_t capacity, const int64_t *store_ids,
                         const int64_t *data_ids, int batch, int first_offset, int n_rows,
                         const void *rows, int64_t row_bytes, int64_t rows_b_stride,
                         const uint8_t *padding_mask, int64_t mask_b_stride, void *stream);

/* ------------------------------------------------------------------------------------
 * SAC update — replaces SAC_Base._train / get_l_probs / _get_td_error for the continuous,
 * stock-network case (sac_base.py:2027-2245; nets: nn_models/q.py:34-91,
 * nn_models/policy.py:116-174, nn_models/layers/linear_layers.py:24-119).
 *
 * A stock net is `depth` ResBlocks (Linear + exact-erf GELU, + input when in == out)
 * followed by a Linear head.  Flat parameter layout per net, all fp32:
 *     W0[H, in] b0[H]  W1[H, H] b1[H] ... W(d-1)[H, H] b(d-1)[H]  Whead[O, H] bhead[O]
 * with O = 1 for a Q net and O = 2A for the policy (rows 0..A-1 mean, A..2A-1 logstd).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    double learning_rate;   /* Adam lr (kept in float64 like the Python float)  */
    int32_t batch;          /* B                                             */
    int32_t seq_len;        /* L = burn_in + n_step + 1                      */
    int32_t burn_in;        /* b                                             */
    int32_t n_step;         /* n  (<= ASAC_MAX_NSTEP)                        */
    int32_t state_size;     /* S                                             */
    int32_t action_size;    /* A  (continuous)                               */
    int32_t ensemble;       /* E  = ensemble_q_num                           */
    int32_t q_hidden, q_depth;
    int32_t pi_hidden, pi_depth;
    int32_t use_n_step_is, use_priority, use_auto_alpha;
    int32_t update_target_per_step;
    int32_t bn_stride;      /* rows per batch element in the action/reward/... arrays (L or L-1) */
    float tau, one_minus_tau /* float32(1. - tau) */, gamma, v_rho, v_c, clip_epsilon, target_c_alpha;
    float td_error_min, td_error_max, per_alpha;
    float gamma_ratio[ASAC_MAX_NSTEP];   /* torch.logspace(0, n-1, n, gamma)    (sac_base.py:285) */
    float lambda_ratio[ASAC_MAX_NSTEP];  /* torch.logspace(0, n-1, n, v_lambda) (sac_base.py:286) */
    int32_t rep_kind;       /* 0: ModelSimpleRep (state = concat obs); 1: trained GRU representation,
                               the states of the online / re-encoded / target representation differ  */
    int32_t rep_param_stride; /* floats of the representation's flat gradient (0 without one): its slice of the
                               peer-exchange buffers (asac_peer_recv_words)                              */
    int32_t ensemble_sample;  /* ensemble_q_sample: the min over the critics runs over the first `ensemble_sample`
                               entries of a random permutation of the members (sac_base.py:1434-1436, 1887);
                               0 or >= ensemble: over all of them                                        */
} AsacSacConfig;

typedef struct {
    /* parameters (flat, see layout above) */
    float *q;            /* [E, Pq]  online critics  (Pq = asac_mlp_param_stride) */
    float *q_target;     /* [E, Pq]  target critics       */
    float *pi;           /* [Ppi]    policy               */
    float *log_alpha;    /* [1]      log_c_alpha          */
    /* Adam moments (torch.optim.Adam defaults, sac_base.py:296-300) */
    float *q_m, *q_v;    /* [E, Pq]  */
    float *pi_m, *pi_v;  /* [Ppi]    */
    float *alpha_m, *alpha_v; /* [1] */
    /* counters: int64[8] = {global_step, adam_step_q, adam_step_pi, adam_step_alpha, adam_step_rep, -, -, -} */
    int64_t *counters;
} AsacSacParams;

typedef struct {
    const float *states;          /* [B, L, S]  encoded states (ModelSimpleRep: concat of vector obs) */
    const float *actions;         /* [B, bn_stride, A] */
    const float *rewards;         /* [B, bn_stride]    */
    const uint8_t *dones;         /* [B, bn_stride]    */
    const uint8_t *last_masks;    /* [B, bn_stride]    */
    const uint8_t *padding_masks; /* [B, bn_stride]    */
    const float *mu_probs;        /* [B, bn_stride, A] */
    const float *priority_is;     /* [B] or NULL       */
    /* injected N(0,1) draws (Normal.rsample / Normal.sample, sac_base.py:1346,1883,1932,2223) */
    const float *eps_y;           /* [B, n+1, A] */
    const float *eps_pi;          /* [B, A]      */
    const float *eps_alpha;       /* [B, A]      */
    const float *eps_td;          /* [B, n+1, A] */
    /* trained representation (cfg.rep_kind != 0; NULL otherwise).  `states` is then the online
     * representation BEFORE its Adam step (seen by _train_rep_q, sac_base.py:2066-2097), */
    const float *states_post;     /* [B, L, S]  online representation re-evaluated after it (:2099-2105):
                                     policy / alpha losses, get_l_probs, Q_i(s_b, a_b) of _get_td_error   */
    const float *target_states;   /* [B, L, S]  target representation (:2073-2078): _get_y inside _get_td_error */
    /* cfg.ensemble_sample < cfg.ensemble: int32[5, E] permutations of the members, one per torch.randperm call of a
     * step in call order — _get_y current rows / next rows (:1434, :1436), _train_policy (:1887), _get_y of
     * _get_td_error current / next rows (asac_ensemble_perms draws them on the device).  NULL: all members. */
    const int32_t *ensemble_perms;
} AsacSacBatch;

typedef struct {
    int32_t n_tiles;      /* ceil(B / asac_sac_tile_batch())                            */
    float *y;             /* [B]       _get_y output of the train pass                  */
    float *tq;            /* [E, B]    target-Q(s0, a0) for the clipped loss            */
    float *q_val;         /* [E, B]    online Q(s0, a0) seen by the loss                */
    float *loss_q;        /* [n_tiles, E] per-tile sums of the weighted loss (÷B = mean) */
    float *grad_q_part;   /* [n_tiles, E, Pq]                                           */
    float *grad_q;        /* [E, Pq]   reduced gradient (also the all-reduce buffer)    */
    float *grad_pi_part;  /* [n_tiles, Ppi]                                             */
    float *grad_pi;       /* [Ppi]                                                      */
    float *stats_pi;      /* [n_tiles, 2] per-tile sums: policy loss, entropy           */
    float *grad_alpha_part; /* [n_tiles, 2] per-tile sums: d loss/d log_alpha, alpha loss */
    float *grad_alpha;    /* [1]                                                        */
    float *pi_probs;      /* [B, L-1, A] get_l_probs output                             */
    float *post_parts;    /* [B, 2+E]  post pass: y' critic part, y' log-prob part, Q_i(s_b,a_b) */
    float *y_td;          /* [B]       _get_y inside _get_td_error                      */
    float *td_error;      /* [B]                                                        */
    float *grad_state;    /* [E, B, S] d(sum_i mean_B loss_i) / d state[:, b] through critic i, or NULL  */
} AsacSacWork;

/* batch elements handled by one CTA of the row-tiled kernels for this configuration
 * (work.n_tiles = ceil(B / asac_sac_tile_batch(cfg))) */
int asac_sac_tile_batch(const AsacSacConfig *cfg);
/* float stride between consecutive flat nets = asac_mlp_param_count rounded up to 4 */
int64_t asac_mlp_param_stride(int in_dim, int hidden, int depth, int out_dim);
/* float count of one flat net: in -> (H x depth) -> out */
int64_t asac_mlp_param_count(int in_dim, int hidden, int depth, int out_dim);

/* 1 when the value pass of this configuration (mode 0: _get_y, mode 1: get_l_probs / _get_td_error) runs on the
 * tcgen05 layer engine (k_value_pass_tc), 0 when it runs on the FFMA row-tile kernel — for reports. */
int asac_sac_value_pass_on_tc(const AsacSacConfig *cfg, int mode);

/* _update_target_variables (sac_base.py:745-764) when counters[0] % update_target_per_step == 0
 * (sac_base.py:2057-2058); `force_tau` >= 0 applies that tau unconditionally (hard copy at
 * start-up, sac_base.py:629). */
int asac_sac_polyak(const AsacSacConfig *cfg, const AsacSacParams *prm, float force_tau, void *stream);

/* _get_y (sac_base.py:1297-1466, continuous branch) + target-Q(s0,a0) (:1541) -> work.y, work.tq */
int asac_sac_target_y(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                      const AsacSacWork *work, void *stream);

/* _train_rep_q's critic part (sac_base.py:1516,1539-1570): forward, clipped loss, backward ->
 * work.q_val, work.loss_q, work.grad_q_part */
int asac_sac_q_backward(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                        const AsacSacWork *work, void *stream);

/* _train_policy (sac_base.py:1882-1908) forward + backward -> work.grad_pi_part, work.stats_pi */
int asac_sac_policy_backward(const AsacSacConfig *cfg, const AsacSacParams *prm,
                             const AsacSacBatch *batch, const AsacSacWork *work, void *stream);

/* _train_alpha's loss (sac_base.py:1930-1945), get_l_probs (:1159-1189) and the network part of
 * _get_td_error (:2182-2245) in one pass over the updated critics / policy ->
 * work.grad_alpha_part, work.pi_probs, work.post_parts.  y' = _get_y is linear in alpha; because
 * the reference evaluates it AFTER the alpha step (:2115-2116 precede :2571) the two linear
 * parts are stored and asac_sac_td_error combines them. */
int asac_sac_post(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                  const AsacSacWork *work, void *stream);

/* which = 0 critics, 1 policy, 2 alpha.  Sums the per-tile partial gradients in tile order
 * into work.grad_*.  (Between this and asac_sac_adam a multi-GPU learner all-reduces
 * work.grad_* and passes grad_scale = 1/world_size.) */
int asac_sac_reduce_grads(const AsacSacConfig *cfg, const AsacSacWork *work, int which, void *stream);

/* torch.optim.Adam.step (betas 0.9/0.999, eps 1e-8) on work.grad_* * grad_scale; advances
 * the optimizer's step counter. */
int asac_sac_adam(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *work,
                  int which, float grad_scale, void *stream);

/* fused single-GPU variant: reduce + Adam in one kernel */
int asac_sac_reduce_adam(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *work,
                         int which, void *stream);

/* _get_td_error's tail (sac_base.py:2223-2245) with the current log_alpha:
 * work.y_td = parts[0] - alpha * parts[1], work.td_error = mean_i |Q_i - y_td|.  (asac_sac_step and
 * asac_sac_reduce_adam(which=2) run it in the same kernel as the alpha Adam step.) */
int asac_sac_td_error(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *work,
                      void *stream);

/* increase_global_step (sac_base.py:2607): counters[0] += 1 */
int asac_sac_advance_step(const AsacSacParams *prm, void *stream);

/* One whole SAC_Base._train + get_l_probs + _get_td_error (sac_base.py:2027-2245,2558-2583):
 * polyak, target_y, q_backward, reduce_adam(q), policy_backward, reduce_adam(pi), post,
 * reduce_adam(alpha), advance_step — the sequence a CUDA graph captures. */
int asac_sac_step(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                  const AsacSacWork *work, void *stream);

/* Data-parallel learner (new capability; the reference trains on one device, sac_base.py:237-243):
 * the gradient exchange runs INSIDE the reduce+Adam kernels over NVLink peer memory.  Every rank
 * allocates one ZERO-INITIALISED receive buffer of asac_peer_recv_words() 8-byte words in memory
 * that all ranks of the node have mapped (torch symmetric memory / cudaIpc); recv[i] is rank i's
 * buffer as seen from THIS process.  A thread pushes {optimizer-step epoch, gradient element} as
 * one 64-bit store into every rank's buffer and polls the same element of all sources until it
 * carries the epoch (flag-in-data, no fences), then sums the sources in rank order: all ranks
 * step with bit-identical mean gradients and no all-reduce kernel sits between the backward
 * pass and Adam.  NULL = single GPU. */
typedef struct {
    int32_t world, rank;
    void *recv[ASAC_MAX_PEERS];
    int64_t recv_words;    /* size of each receive buffer in 8-byte words */
} AsacPeerTable;
int64_t asac_peer_recv_words(const AsacSacConfig *cfg, int world);
/* The polls above are bounded (20 s): a crashed peer, or ranks that called train() a different number of
 * times, would otherwise hang the GPU inside a kernel.  An abandoned wait counts the missing gradient as
 * zero and bumps a device counter; this returns it (and clears it when reset != 0).  Synchronises with the
 * device.  The learner polls it every 256 steps and on close() and raises (no reference counterpart: the
 * reference trains on one device). */
int asac_peer_timeouts(int reset);

/* asac_sac_step without its tail: [polyak,] target_y, q_backward, reduce_adam(q), policy_backward,
 * reduce_adam(pi), post.  `with_polyak` = 0 when the caller has enqueued asac_sac_polyak itself
 * (e.g. on a parallel graph branch).  `peers` != NULL: gradients are averaged over the ranks.
 * Same results as the stand-alone calls in that order, bit for bit; what this entry point adds is knowledge of
 * the sequence: each kernel runs its predecessor-independent prologue before griddepcontrol.wait, the train pass
 * leaves the policy's pre-activations for the policy backward (scratch: the tile slices of work->grad_pi_part)
 * and the policy backward leaves Q_i(s_b, a_b) for the post pass (scratch: work->tq, which therefore holds the
 * ONLINE critics' values after the call). */
int asac_sac_step_networks(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                           const AsacSacWork *work, int with_polyak, const AsacPeerTable *peers,
                           void *stream);

/* The tail of SAC_Base.train() for a prioritized run in one single-CTA kernel (batch <= 1024):
 * alpha reduce + Adam (sac_base.py:1941-1948), _get_td_error's tail with the updated alpha
 * (:2223-2245) -> work.y_td / work.td_error, PrioritizedReplayBuffer.update on those td errors
 * (replay_buffer.py:412-427; cfg->td_error_min/max/per_alpha), and the global-step / optimizer
 * counters (sac_base.py:2607).  Equals asac_sac_step's tail followed by asac_per_update.
 * nodes == NULL defers the tree update: td errors are written, the caller applies them later with
 * asac_per_update(work->td_error) (the learner does so on the sample-ahead branch of its next step). */
int asac_sac_finish_step(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *work,
                         float *nodes, int64_t capacity, const int64_t *store_ids,
                         const int64_t *data_ids, double *per_state, const AsacPeerTable *peers,
                         void *stream);

/* ------------------------------------------------------------------------------------
 * Recurrent representation — replaces the stock multi-layer GRU wrapper
 * (algorithm/nn_models/layers/seq_layers.py:14-114, no padding mask) inside a plugin ModelRep of the
 * form of envs/test/nn_rnn.py:6-21: state, hn = GRU(cat[obs, pre_action], h0), called by
 * get_l_states (sac_base.py:1118-1146) three times per _train (:2066-2105), trained by the critic
 * loss through every burn-in step (:1573-1601) and by choose_action on the actor side (:1003).
 *
 * Flat parameter layout, layer after layer, torch.nn.GRU's own order and gate order (r, z, n):
 *     weight_ih[3H, in_l]  weight_hh[3H, H]  bias_ih[3H]  bias_hh[3H]      in_0 = obs + action, in_l = H
 * Cell, as ATen evaluates it: r = sigmoid(Wir x + bir + Whr h + bhr), z likewise,
 * n = tanh(Win x + bin + r * (Whn h + bhn)), h' = (h - n) * z + n.
 * ---------------------------------------------------------------------------------- */
#define ASAC_GRU_MAX_LAYERS 4
typedef struct {
    int32_t obs_size;    /* floats of the (vector) observation per step        */
    int32_t action_size; /* floats of pre_action appended to it                */
    int32_t hidden;      /* H (<= 64); the state is the top layer's output     */
    int32_t layers;      /* <= ASAC_GRU_MAX_LAYERS; hidden state shape (layers, H) */
} AsacGruShape;

int64_t asac_gru_param_count(const AsacGruShape *shape);
/* sequences one CTA of asac_gru_backward handles (>= 1 when t_grad + 1 steps fit in shared memory) */
int asac_gru_backward_tile(const AsacGruShape *shape, int t_grad);

/* GRU.forward over [batch, seq_len]: x_t = [obs[b, t], pre_action], pre_action = pre_actions[b, t] when
 * given, else actions[b, t-1] (zeros at t = 0): gen_n_pre_actions(keep_last_action=True),
 * utils/operators.py:39-59.  h0: [batch] rows of [layers, H], h0_b_stride floats apart (NULL = zeros).
 * Outputs: states [batch, seq_len, H]; hn [batch, seq_len, layers, H] (may be NULL);
 * save [batch, seq_len, layers, 4H] = (r, z, n, Whn h + bhn) for asac_gru_backward (may be NULL).
 * n_nets (1 or 2) parameter sets run in one launch on the same inputs (online + target). */
typedef struct {
    const float *params;
    float *states, *hn, *save;
} AsacGruNet;
int asac_gru_forward(const AsacGruShape *shape, const AsacGruNet *nets_host, int n_nets, const float *obs,
                     const float *actions, int bn_stride, const float *pre_actions, const float *h0,
                     int64_t h0_b_stride, int batch, int seq_len, void *stream);

/* Back-propagation through time of sum_e grad_state[e, b, :] applied to the top layer's output at
 * step t_grad (= burn_in: the only state the critic loss reads, sac_base.py:1510) back to step 0,
 * through every layer.  hn / save come from asac_gru_forward with the same parameters and inputs.
 * grad_part[batch, P rounded up to 4]: one partial gradient per sequence, summed in sequence order by
 * asac_flat_reduce_adam. */
int asac_gru_backward(const AsacGruShape *shape, const float *params, const float *obs, const float *actions,
                      int bn_stride, const float *pre_actions, const float *h0, int64_t h0_b_stride, int batch,
                      int seq_len, int t_grad, const float *grad_state, int ensemble, const float *hn,
                      const float *save, float *grad_part, void *stream);

/* Deterministic sum of n_tiles partial gradients (tile_stride floats apart) -> grad[count], then
 * torch.optim.Adam.step on param with moments m, v (the kernel of asac_sac_reduce_adam on a caller-
 * supplied flat buffer).  step_counter[0] = steps taken so far; NOT advanced by this call. */
int asac_flat_reduce_adam(float *param, float *m, float *v, const float *grad_part, int n_tiles,
                          int64_t tile_stride, int64_t count, float *grad, const int64_t *step_counter,
                          double learning_rate, void *stream);
/* target = target * one_minus_tau + source * tau when counters[0] % per_step == 0 or force != 0
 * (_update_target_variables, sac_base.py:745-764, for the representation's parameters) */
int asac_flat_polyak(float *target, const float *source, int64_t count, const int64_t *counters, int per_step,
                     float tau, float one_minus_tau, int force, void *stream);

/* Everything the learner holds for a GRU representation (device pointers owned by the caller). */
typedef struct {
    AsacGruShape shape;
    float *params, *params_target, *m, *v;  /* [P] each                                             */
    const float *obs;          /* [B, L, obs_size]   gathered observation windows                   */
    const float *h0;           /* first row of the gathered pre_seq_hidden_state windows            */
    int64_t h0_b_stride;       /* floats between batch elements of h0 (L * layers * H)              */
    float *states;             /* [B, L, H]  out: online representation before its step             */
    float *states_post;        /* [B, L, H]  out: online representation after its step              */
    float *target_states;      /* [B, L, H]  out: target representation                             */
    float *hn;                 /* [B, L, layers, H]  scratch: hidden states of the first pass       */
    float *hn_post;            /* [B, L, layers, H]  out: next_bnx_seq_hidden_states (:2099-2105)   */
    float *save;               /* [B, L, layers, 4H] scratch: gates of the first pass               */
    float *grad_part;          /* [rep_tiles, P rounded up to 4]                                    */
    float *grad;               /* [P]  reduced gradient                                             */
    int32_t rep_tiles;         /* rows of grad_part = B (one per sequence)                          */
    int32_t reserved_;
} AsacGruRep;

/* asac_sac_step_networks for a run with a trained GRU representation (sac_base.py:2057-2116):
 * [polyak of critics and representation,] online + target representation, _get_y, critic loss and
 * backward (+ d loss / d state), critic Adam, representation BPTT + Adam, representation again,
 * policy loss / Adam, post pass on (states_post, target_states).  batch->states / states_post /
 * target_states and work->grad_state must be the buffers named in `rep`.  `peers` != NULL: the three
 * gradients (critics, representation, policy) are averaged over the ranks inside their reduce+Adam
 * kernels, as in asac_sac_step_networks.  Follow with asac_sac_finish_step (advances counters[4] too when
 * cfg->rep_kind != 0) or, single GPU, asac_sac_staged_tail. */
int asac_sac_step_networks_rep(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *batch,
                               const AsacSacWork *work, const AsacGruRep *rep, int with_polyak,
                               const AsacPeerTable *peers, void *stream);

/* asac_sac_step's tail on its own: alpha reduce + Adam, td error, all step counters
 * (the non-prioritized / batch > 1024 counterpart of asac_sac_finish_step). */
int asac_sac_staged_tail(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *work, void *stream);

/* N(0,1) draws for eps_* (Philox4x32-10 + Box-Muller), keyed by (seed, counter[0], stream_id) */
int asac_fill_normal(float *out, int64_t n, uint64_t seed, const int64_t *counter, int stream_id,
                     void *stream);

/* Pulls up to 8 device regions into L2 (prefetch.global.L2, one 128-byte line per thread): the learner
 * issues it on the side branch of a step for the buffers the critical-path kernels will touch first
 * (policy parameters, Adam moments) while the sample / gather kernels run. */
int asac_l2_prefetch(const void *const *regions_host, const int64_t *bytes_host, int n_regions, void *stream);

/* Stock-net forward for the actor side / tests: out[rows, O] = net(x[rows, in]). */
int asac_mlp_forward(const float *params, int in_dim, int hidden, int depth, int out_dim,
                     const float *x, int64_t rows, float *out, void *stream);

/* The same forward on the tcgen05 tensor cores (kind::tf32, accumulator in TMEM, 3xTF32 split
 * operands so that results stay within 1e-5 of the fp32 reference): one persistent CTA per SM over
 * 128-row tiles.  hidden == 64 (the stock width), out_dim <= 16.  Meant for large row counts (actor-side
 * batches, large-batch value passes); asac_mlp_forward stays the exact-fp32 FFMA path. */
int asac_mlp_forward_tc(const float *params, int in_dim, int hidden, int depth, int out_dim,
                        const float *x, int64_t rows, float *out, void *stream);

/* The layer engine of the UPDATE kernels on its own (tests, actor side): the same forward with the
 * tensor-core tile turned around — D[64 features, R rows] = W . X^T, UMMA_M = hidden width (64), UMMA_N =
 * rows_per_cta (multiple of 8, 8..256) — so that small row tiles (16 rows of a batch-256 step) and long
 * windows (208 rows of get_l_probs over a burn-in) are both ONE batch of tcgen05.mma per layer.
 * variant 0 / 1 selects the TMEM fragment shape of the epilogue (16x256b / 32x32b); both must agree. */
int asac_mlp_forward_tcf(const float *params, int in_dim, int hidden, int depth, int out_dim,
                         const float *x, int64_t rows, float *out, int rows_per_cta, int variant, void *stream);
/* debug: same launch; probe[(depth + 1) * 5] (device int64) receives, per layer, CTA 0's clocks at
 * {issue start, MMAs issued, MMAs complete, epilogue done, layer done} */
int asac_mlp_forward_tcf_probe(const float *params, int in_dim, int hidden, int depth, int out_dim,
                               const float *x, int64_t rows, float *out, int rows_per_cta, int variant,
                               long long *probe, void *stream);

/* Actor side: SAC_Base._choose_action for a stock continuous policy (sac_base.py:882-966, branch
 * :943-964): policy forward on `states` [rows, S] (asac_mlp_forward, or the tcgen05 kernel when
 * use_tensor_cores != 0), then c_action = offline_action | tanh(mean) | tanh(Normal.sample()) and
 * prob = squash_correction_prob(policy, atanh(clamp(c_action, +-0.999))) (utils/operators.py:17-19).
 * eps: injected N(0,1) draws [rows, A] (NULL: Philox keyed by seed, counter[0]); scratch: [rows, 2A]. */
int asac_policy_act(const float *params, int state_size, int hidden, int depth, int action_size,
                    const float *states, int64_t rows, const float *eps, const float *offline_action,
                    int disable_sample, uint64_t seed, const int64_t *counter, float *scratch,
                    float *out_action, float *out_prob, int use_tensor_cores, void *stream);

/* Debug aid: SM clock stamps (clock64) taken by CTA (0,0) at the phase boundaries of the last value
 * pass [0], critic backward [1] and policy backward [2] launch; out_host is int64[3][32], slot 31 is
 * the kernel's exit (tools/phase_breakdown.py prints the differences).  Synchronises. */
int asac_debug_phase_clocks(int64_t *out_host);
/* debug (library built with -DASAC_PROBES): 8 %globaltimer stamps (ns) around the critics' Adam step, then reset */
int asac_debug_global_stamps(uint64_t *out_host);

/* ------------------------------------------------------------------------------------
 * Discrete (and the discrete half of hybrid) action branches — replaces the d_action_sizes paths of
 * SAC_Base._get_y / _train_rep_q / _train_policy / _train_alpha / get_l_probs / _get_td_error
 * (sac_base.py:1356-1421, 1543-1570, 1858-1880, 1924-1929, 1176-1178, 2226-2230) for the stock nets: one
 * LinearLayers(state -> d_dense_n x d_dense_depth -> d_action_size_k) per branch in every ModelQ and in
 * ModelPolicy (nn_models/q.py:60-64, policy.py:143-147).  Flat layout of one member: the K branch nets one
 * after the other, each in the stock layout (W0 b0 ... Whead bhead) padded to a multiple of 4 floats.
 * Actions and mu-probabilities are stored as [one-hot per branch ..., continuous ...] rows of D + A floats.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t branches;
    int32_t sizes[ASAC_MAX_BRANCHES];   /* d_action_sizes                                                    */
    int32_t hidden, depth;              /* d_dense_n, d_dense_depth                                          */
    int32_t state_size;
    float target_d_alpha;               /* ratio; the per-column target is ratio * -log(1 / size) (:463-466) */
    float entropy_penalty;              /* d_policy_entropy_penalty                                          */
} AsacDiscreteConfig;

int64_t asac_dnets_member_floats(const AsacDiscreteConfig *d);   /* floats of one member's K branch nets          */
int asac_dnets_tiles(int rows);                                  /* 16-row tiles = partial-gradient rows          */
/* out[member, row, column] = branch nets of every member on row r = x + r * x_row_stride  (forward only) */
int asac_dnets_forward(const AsacDiscreteConfig *d, const float *params, int64_t member_stride, int members,
                       const float *x, int64_t x_row_stride, int rows, float *out, void *stream);
/* the same walk with saved activations, then the backward pass from d_out[member, row, column] = d loss / d output;
 * grad_part[tile, member, member_floats]: partial gradients per 16-row tile (summed by asac_flat_reduce_adam);
 * d_x (may be NULL) [member, branch, row, state_size]: d loss / d input row, for a trained representation */
int asac_dnets_backward(const AsacDiscreteConfig *d, const float *params, int64_t member_stride, int members,
                        const float *x, int64_t x_row_stride, int rows, const float *d_out, float *grad_part,
                        float *d_x, void *stream);
/* d_y[B]: expectation of mean-ensemble Q minus alpha log pi under the policy on rows b..b+n, V-trace with the
 * discrete importance ratio (sac_base.py:1384-1412, 1244-1295).  pi_logits [B * L, D], tq [E, B * L, D];
 * pi_probs_d != NULL: the td-error pass, whose mu probabilities are the policy's own (:2233). */
int asac_d_target(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *pi_logits, const float *tq,
                  const float *actions_full, const float *mu_full, const float *pi_probs_d, const float *rewards,
                  const uint8_t *dones, const uint8_t *last_masks, const uint8_t *padding_masks,
                  const float *log_d_alpha, float *d_y, void *stream);
/* discrete_dqn_like: d_y of get_dqn_like_d_y (sac_base.py:1193-1242, 1363-1383) — double DQN on the last solid step.
 * eval_q / tq [E, B * L, D]: online / target critics on every row; perm_target / perm_online int32[E]: the
 * reference's two torch.randperm draws (member j of one shuffled stack is paired with member j of the other), NULL:
 * identity. */
int asac_d_target_dqn(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *eval_q, const float *tq,
                      const int32_t *perm_target, const int32_t *perm_online, const float *rewards,
                      const uint8_t *dones, const uint8_t *last_masks, const uint8_t *padding_masks, float *d_y,
                      void *stream);
/* critic loss of the discrete part and its gradient w.r.t. the critics' outputs (:1543-1547, 1564-1570) */
int asac_d_q_grad(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *q, const float *actions_full,
                  const float *d_y, const float *weights, float scale, float *d_out, float *loss, float *q_single,
                  void *stream);
/* policy loss incl. the entropy penalty and its gradient w.r.t. the logits (:1858-1880); the logits of batch
 * element e are row e * stride_rows + row_off of `logits` */
int asac_d_pi_grad(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, int stride_rows,
                   int row_off, const float *q, const float *mu_full, const float *log_d_alpha, float *d_out,
                   float *loss, float *entropy, void *stream);
/* get_l_probs, discrete columns (:1176-1178): pi_probs_d [B, L - 1, D] (+ columns [0, D) of pi_probs_full) */
int asac_d_probs(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, float *pi_probs_d,
                 float *pi_probs_full, void *stream);
/* _train_alpha, discrete part (:1924-1929) + Adam on log_d_alpha; step[0] = steps taken so far, not advanced */
int asac_d_alpha(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, float *log_d_alpha,
                 float *m, float *v, const int64_t *step, float grad_scale, float *grad_out, float *loss_out,
                 void *stream);
/* td_error[e] (+)= mean_i |sum(onehot * q_i) / branches - d_y|  (:2226-2230, 2238-2243) */
int asac_d_td(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *q, const float *actions_full,
              const float *d_y, float *td_error, int accumulate, void *stream);
/* out[n_perms, E] = independent uniform random permutations of 0..E-1 (Fisher-Yates on Philox4x32-10 keyed by
 * seed, counter[0] and the permutation's index): the torch.randperm(E) draws of a step */
int asac_ensemble_perms(int32_t *out, int n_perms, int ensemble, uint64_t seed, const int64_t *counter, void *stream);
/* counters[i] += 1 for every bit i of mask (0 global step, 1 critics, 2 policy, 3 alpha, 4 representation) */
int asac_bump_counters(int64_t *counters, int mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ASAC_B200_H */
