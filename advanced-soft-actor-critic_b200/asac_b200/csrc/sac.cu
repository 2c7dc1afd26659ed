// Fused SAC update kernels (sm_100a): value pass (_get_y / get_l_probs / _get_td_error /
// alpha loss), critic forward+loss+backward, policy forward+backward, gradient reduction,
// Adam and the Polyak target update.
//
// Replaces the torch op graph of SAC_Base._train and friends for the continuous-action,
// stock-network case (algorithm/sac_base.py:745-764, 1159-1189, 1244-1466, 1468-1605,
// 1841-1949, 2027-2126, 2182-2245).  Exact formulas and the reference quirks that are
// reproduced on purpose are listed in DESIGN.md §5.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "mlp_tile.cuh"
#include "tc_engine.cuh"
#include "tree_apply.cuh"

namespace cg = cooperative_groups;

namespace asac {

struct SacArgs {
    // 4-D TMA tensor maps {k, row, layer, member} over the hidden -> hidden layers of the flat parameter
    // buffers: [0] online critics, [1] target critics, [2] policy (valid when use_tma != 0)
    alignas(64) CUtensorMap maps[3];
    int use_tma;
    AsacSacConfig cfg;
    AsacSacParams prm;
    AsacSacBatch bat;
    AsacSacWork wrk;
    int tile_batch;  // batch elements per CTA
    int mode;        // value pass: 0 = train (_get_y), 1 = post (alpha loss, l_probs, td error)
    int late_wait;   // 1 inside the fused step chains (asac_sac_step*): the kernels know what runs ahead of them and
                     // put griddepcontrol.wait behind their predecessor-independent prologue; 0 for the stand-alone
                     // entry points, whose predecessor is whatever the caller launched: wait first.  ASAC_LATE_WAIT=0
                     // keeps every wait at the top (experiments)
    int push_combine;  // value pass: ranks 1.. push their value rows to rank 0 (ASAC_PUSH_COMBINE=0: rank 0 pulls them)
    int pi_handoff;    // fused step, no trained representation, one policy pass per tile: the TRAIN value pass (rank 0)
                       // leaves the policy's pre-activations of all its rows in the tile's slice of wrk.grad_pi_part
                       // ([depth][16][lda]; the partial gradients are written there only at the end of the policy
                       // backward) and the policy backward rebuilds its saved forward from the s_b rows instead of
                       // running the trunk again — the policy does not change in between
    int q_sb_handoff;  // fused step: the policy backward also evaluates Q_i(s_b, a_b) (free rows of its critic pass,
                       // the critics' weights no longer change within the step) and leaves it in wrk.tq; the post
                       // pass reads it there instead of running the online critics itself (sac_base.py:2211-2216)
    int plan[24];    // the launching kernel's shared-memory plan (ValuePlan / GradPlan), computed on the host: every
                     // thread re-deriving it cost ~150 instructions with two integer divisions at kernel entry
};

__host__ __device__ __forceinline__ NetShape q_shape(const AsacSacConfig &c) {
    return NetShape{c.state_size + c.action_size, c.q_hidden, c.q_depth, 1};
}
__host__ __device__ __forceinline__ NetShape pi_shape(const AsacSacConfig &c) {
    return NetShape{c.state_size, c.pi_hidden, c.pi_depth, 2 * c.action_size};
}
__host__ __device__ __forceinline__ int sac_lda(const AsacSacConfig &c) {
    const int a = tile_lda(c.q_hidden, c.state_size + c.action_size), b = tile_lda(c.pi_hidden, c.state_size);
    return a > b ? a : b;
}
__host__ __device__ __forceinline__ int sac_wsz(const AsacSacConfig &c) {
    const int a = tile_wsz(c.q_hidden, c.state_size + c.action_size), b = tile_wsz(c.pi_hidden, c.state_size);
    return round_up(a > b ? a : b, 256);
}
// first 1024-byte aligned address at or after sm + off (the plans reserve 256 floats of slack for it)
__device__ __forceinline__ float *aligned_slots(float *sm, int off) {
    const unsigned a = smem_u32(sm + off);
    return sm + off + (((1024u - (a & 1023u)) & 1023u) >> 2);
}
__host__ __device__ __forceinline__ int sac_part(const AsacSacConfig &c) {
    return tile_part_floats(c.q_hidden > c.pi_hidden ? c.q_hidden : c.pi_hidden);
}
// floats reserved in front of the weight slots for the job table and the slot mbarriers
constexpr int PIPE_HEADER_FLOATS = (MAX_WEIGHT_JOBS * (int)sizeof(WeightJob) + MAX_WEIGHT_SLOTS * 8 + 15) / 16 * 4;
constexpr int kSmemBudgetFloats = 227 * 1024 / 4;
// weight slots that fit behind `fixed` floats of other shared memory (at most n_jobs; a single slot —
// hidden width 128 in the backward kernels — still works, the next layer is then staged only after
// the current one has been consumed)
__host__ __device__ __forceinline__ int slots_that_fit(int fixed, int slot_floats, int n_jobs) {
    int n = (kSmemBudgetFloats - fixed) / slot_floats;
    if (n > n_jobs) n = n_jobs;
    if (n > MAX_WEIGHT_SLOTS) n = MAX_WEIGHT_SLOTS;
    return n < 1 ? 1 : n;
}

// Phase clocks of CTA (0,0) (debug aid read back by asac_debug_phase_clocks): kernel 0 = value pass,
// 1 = critic backward, 2 = policy backward; slot 31 = kernel exit.
__device__ long long g_phase_clock[3][32];
#define ASAC_PHASE(k, i)                                                              \
    do {                                                                              \
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_phase_clock[k][i] = clock64(); \
    } while (0)

// Kernel parameters live in a constant bank that every SM faults in line by line on first use: the ~1.2 KB of
// SacArgs read field after field by the setup code cost ~2 us of serial misses at the top of every kernel (measured
// with the phase clocks).  Each warp touches a different 64-byte line first, so the misses overlap.
__device__ __forceinline__ void warm_kernel_params(const SacArgs &a) {
    const int *words = reinterpret_cast<const int *>(&a);
    int acc = 0;
    for (int off = (threadIdx.x >> 5) * 16; off < (int)(sizeof(SacArgs) / 4); off += (NT / 32) * 16) acc ^= words[off];
    asm volatile("" ::"r"(acc));
}

// debug: %globaltimer stamps (ns) comparable across SMs — [0] critic backward exit (max over CTAs), [4] first / [5]
// last policy-backward CTA start, [6] last return of its griddepcontrol.wait (tools/pdl_overlap_probe.py)
__device__ unsigned long long g_gt[8];
__device__ __forceinline__ void gt_stamp(int i, bool take_min) {
#ifdef ASAC_PROBES
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (take_min) atomicMin(&g_gt[i], t); else atomicMax(&g_gt[i], t);
    }
#endif
}

constexpr float LOG_SQRT_2PI = 0.91893853320467274178f;

// torch.distributions.Normal.log_prob: -((x - loc)^2) / (2 var) - log(scale) - log(sqrt(2 pi))
__device__ __noinline__ float normal_log_prob(float x, float loc, float scale) {
    const float var = scale * scale;
    const float d = x - loc;
    return -(d * d) / (2.f * var) - logf(scale) - LOG_SQRT_2PI;
}
// max(1 - tanh(x)^2, 1e-2)   (utils/operators.py:14,19)
__device__ __noinline__ float squash_floor(float x) {
    const float t = tanhf(x);
    return fmaxf(1.f - t * t, 1e-2f);
}
// policy.py:169
__device__ __noinline__ float policy_loc(float m) { return tanhf(m / 5.f) * 5.f; }
__device__ __noinline__ float policy_scale(float s) { return expf(fminf(fmaxf(s, -20.f), 0.5f)); }

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    return s;
}

// ensemble_q_sample < ensemble_q_num: the min over the critics runs over a random subset (sac_base.py:1434-1436, 1887)
__host__ __device__ __forceinline__ bool ensemble_subset(const AsacSacConfig &c) {
    return c.ensemble_sample > 0 && c.ensemble_sample < c.ensemble;
}
// Rank 0 of a value-pass cluster: min over the members of every value row.  All members, or — with a subset —
// qmin[r] over the first Es entries of `perm_cur` (the rows used as V_k) and qmin2[r] over those of `perm_next`
// (the rows used as V_{k+1}); the reference draws the two subsets independently.
template <typename Cluster>
__device__ __forceinline__ void combine_value_rows(Cluster &cluster, const AsacSacConfig &c, const int32_t *perms,
                                                   int first_perm, float *qmin, float *qmin2, int RV) {
    const int E = c.ensemble, tid = threadIdx.x;
    if (!ensemble_subset(c) || perms == nullptr) {
        for (int i = 1; i < E; ++i) {
            const float *rmin = cluster.map_shared_rank(qmin, i);
            for (int r = tid; r < RV; r += NT) qmin[r] = fminf(qmin[r], rmin[r]);  // sac_base.py:1439-1442
        }
        return;
    }
    const int Es = c.ensemble_sample;
    const int32_t *pc = perms + first_perm * E, *pn = pc + E;
    for (int r = tid; r < RV; r += NT) {
        float m1 = INFINITY, m2 = INFINITY;
        for (int j = 0; j < Es; ++j) {
            m1 = fminf(m1, cluster.map_shared_rank(qmin, pc[j])[r]);
            m2 = fminf(m2, cluster.map_shared_rank(qmin, pn[j])[r]);
        }
        qmin2[r] = m2;
        qmin[r] = m1;  // (own row r is read above by this thread only)
    }
}

// ------------------------------------------------------------------------------------ smem plans
struct ValuePlan {
    int lda, wsz, rows_max;  // rows_max: multiple of 16
    int off_xin, off_a, off_b, off_ho, off_xs, off_logp, off_qmin, off_qpush, off_qmin2, off_ratio, off_qs, off_red, off_part,
        off_heads, off_pipe, off_slots;
    int n_jobs, n_slots;
    int total;  // floats
};
__host__ __device__ __forceinline__ ValuePlan value_plan(const AsacSacConfig &c, int TB, int mode, int q_sb_handoff = 0) {
    ValuePlan p;
    const int L = c.seq_len, n = c.n_step, A = c.action_size;
    const int t0 = (mode == 1 && c.use_n_step_is) ? 0 : c.burn_in;
    const int Lp = L - t0;
    // post pass of a run with a trained representation: the value rows go through the policy a second
    // time on the TARGET representation's states (sac_base.py:2571-2582)
    const int extra = (mode == 1 && c.rep_kind != 0) ? TB * (n + 1) : 0;
    const int rp = round_up(TB * Lp + extra, PASS_ROWS);
    const int rq = round_up(TB * (n + 1), PASS_ROWS) + round_up(TB, PASS_ROWS);
    p.rows_max = rp > rq ? rp : rq;
    p.lda = sac_lda(c);
    p.wsz = sac_wsz(c);
    int o = 0;
    p.off_xin = o; o += p.rows_max * p.lda;
    p.off_a = o; o += p.rows_max * p.lda;
    p.off_b = o; o += p.rows_max * p.lda;
    p.off_ho = o; o += round_up(rp * 2 * A, 4);
    p.off_xs = o; o += round_up(TB * (n + 1) * A, 4);
    p.off_logp = o; o += round_up(TB * (n + 1), 4);
    p.off_qmin = o; o += round_up(p.rows_max, 4);
    p.off_qpush = o; o += (c.ensemble - 1) * round_up(p.rows_max, 4);  // ranks 1.. push their value rows to rank 0
    p.off_qmin2 = o; o += ensemble_subset(c) ? round_up(p.rows_max, 4) : 0;  // min over the "next rows" subset
    p.off_ratio = o; o += round_up(TB * (n > 0 ? n : 1), 4);
    p.off_qs = o; o += round_up(c.ensemble * TB, 4);
    p.off_red = o; o += 32;
    p.off_part = o; o += sac_part(c);
    p.off_heads = o; o += head_floats(c.pi_hidden, 2 * A) + 2 * head_floats(c.q_hidden, 1);  // policy, target, online
    p.off_pipe = o; o += PIPE_HEADER_FLOATS;
    p.off_slots = o;
    o += 256;  // the slots start on the next 1024-byte boundary (TMA 128-byte swizzle)
    p.n_jobs = c.pi_depth + c.q_depth + ((mode == 1 && !q_sb_handoff) ? c.q_depth : 0);
    p.n_slots = slots_that_fit(o, p.wsz, p.n_jobs);
    o += p.n_slots * p.wsz;
    p.total = o;
    return p;
}

struct GradPlan {
    int lda, wsz;
    int off_px, off_pz, off_qz, off_qin, off_g0, off_g1, off_g2, off_small, off_red, off_part, off_heads, off_pipe,
        off_slots;
    int n_jobs, n_slots;
    int total;
};
// critic: px/pz hold the critic's activations, qz unused.  policy: px/pz policy, qz critics' z.
__host__ __device__ __forceinline__ GradPlan grad_plan(const AsacSacConfig &c, bool policy, int pi_handoff = 0) {
    GradPlan p;
    p.lda = sac_lda(c);
    p.wsz = sac_wsz(c);
    const int rows = PASS_ROWS;
    const int d = policy ? c.pi_depth : c.q_depth;
    int o = 0;
    p.off_px = o; o += (d + 1) * rows * p.lda;
    p.off_pz = o; o += d * rows * p.lda;
    p.off_qz = o; o += policy ? c.q_depth * rows * p.lda : 0;  // this CTA's critic only (cluster of E CTAs)
    p.off_qin = o; o += policy ? rows * p.lda : 0;
    p.off_g0 = o; o += rows * p.lda;
    p.off_g1 = o; o += rows * p.lda;
    p.off_g2 = o; o += rows * p.lda;
    p.off_small = o;  // (policy: + 3 A per row of loss terms, + (E - 1) A per row of action gradients pushed by the other ranks)
    o += round_up(rows * ((9 + c.ensemble - 1) * c.action_size + c.ensemble + 4), 4);
    p.off_red = o; o += 32;
    p.off_part = o; o += sac_part(c);
    p.off_heads = o; o += head_floats(c.pi_hidden, 2 * c.action_size) + head_floats(c.q_hidden, 1);  // policy, critic
    p.off_pipe = o; o += PIPE_HEADER_FLOATS;
    p.off_slots = o;
    // critic kernel: forward + reverse walk of one critic; policy kernel: policy forward, critic forward,
    // critic reverse, policy reverse (the last only on cluster rank 0, the plan reserves for it anyway)
    o += 256;  // the slots start on the next 1024-byte boundary (TMA 128-byte swizzle)
    p.n_jobs = policy ? ((pi_handoff ? 0 : c.pi_depth) + c.q_depth + (c.q_depth - 1) + (c.pi_depth - 1))
                      : (c.q_depth + c.q_depth - 1);
    p.n_slots = slots_that_fit(o, p.wsz, p.n_jobs);
    o += p.n_slots * p.wsz;
    p.total = o;
    return p;
}

// ------------------------------------------------------------------------------------ value pass
// mode 0: y = _get_y(...) on the online policy / target critics (sac_base.py:1297-1466) and
//         tq_i = target-Q_i(s_b, a_b) (sac_base.py:1541).
// mode 1: after the three Adam steps — alpha-loss terms (:1930-1945), pi_probs = get_l_probs
//         (:1159-1189), y' = _get_y with mu := pi_probs and td = mean_i |Q_i(s_b,a_b) - y'|
//         (:2182-2245).
// grid (n_tiles, E), thread-block cluster (1, E, 1): the E CTAs of a batch tile each run the
// policy (duplicated, it is the cheap part) and ONE ensemble member; the member outputs are
// combined by cluster rank 0 through distributed shared memory (min over i, sac_base.py:1439-1442).
__global__ void __launch_bounds__(NT) k_value_pass(const __grid_constant__ SacArgs a) {
    // Programmatic dependent launch: the post pass follows the policy's Adam step, which only touches the policy's
    // parameters — the job table, the critics' heads and the policy rows' states are staged while it drains and
    // griddepcontrol.wait sits in front of the first read of the policy.  (Also with a trained representation: the
    // states of the post pass were written before the policy backward, two kernels ahead of that optimiser step.)
    // The TRAIN pass waits first: with a representation its states come from the kernel just ahead.
    const bool late_wait = a.mode == 1 && a.late_wait;
    if (!late_wait) pdl_wait();
    pdl_trigger();
    warm_kernel_params(a);
    cg::cluster_group cluster = cg::this_cluster();
    const int net = (int)cluster.block_rank();
    cluster.barrier_arrive();  // matched by barrier_wait() in front of the first remote access: every rank is running
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const AsacSacConfig &c = a.cfg;
    const int tid = threadIdx.x;
    const int B = c.batch, L = c.seq_len, b = c.burn_in, n = c.n_step, S = c.state_size, A = c.action_size;
    const int E = c.ensemble, TB = a.tile_batch;
    const bool post = a.mode == 1;
    const bool use_is = c.use_n_step_is != 0;
    const int e0 = blockIdx.x * TB;
    const int TBa = min(TB, B - e0);
    const int t0 = (post && use_is) ? 0 : b;
    const int Lp = L - t0;
    const int RP = TBa * Lp, RV = TBa * (n + 1), RS = TBa;
    const bool need_tq = !post && c.clip_epsilon > 0.f;
    const bool online = post && !a.q_sb_handoff;  // the online critics' Q_i(s_b, a_b) is computed here
    // which representation's states feed what (sac_base.py:2066-2105, 2558-2582): the train pass sees the
    // online states before the representation's Adam step everywhere; the post pass evaluates the
    // probabilities, the alpha loss and Q_i(s_b, a_b) on the re-encoded online states (st_p) and
    // _get_y on the target representation's states (st_v).  Without a trained representation all three coincide.
    const float *st_p = post && a.bat.states_post ? a.bat.states_post : a.bat.states;
    const float *st_v = post && a.bat.target_states ? a.bat.target_states : st_p;
    const bool split = post && c.rep_kind != 0;  // value rows run through the policy separately (rows RP ...)
    const int RPt = RP + (split ? RV : 0);
    const ValuePlan &pl = *reinterpret_cast<const ValuePlan *>(a.plan);
    const int lda = pl.lda;
    float *xin = sm + pl.off_xin, *bufA = sm + pl.off_a, *bufB = sm + pl.off_b;
    float *ho = sm + pl.off_ho, *xs = sm + pl.off_xs, *logp = sm + pl.off_logp, *qmin = sm + pl.off_qmin;
    float *ratio = sm + pl.off_ratio, *qs = sm + pl.off_qs, *red = sm + pl.off_red, *part = sm + pl.off_part;

    const NetShape ps = pi_shape(c), qsh = q_shape(c);
    const int64_t q_stride = net_stride(qsh);

    ASAC_PHASE(0, 0);
    // ---- weight pipe: policy trunk, target critic `net`, (post) online critic `net`
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    if (tid < 32) {  // one lane per job
        const CUtensorMap *mq = a.use_tma ? &a.maps[0] : nullptr, *mt = a.use_tma ? &a.maps[1] : nullptr,
                          *mp = a.use_tma ? &a.maps[2] : nullptr;
        if (a.use_tma && tid < 3 && (tid != 0 || online)) prefetch_tensormap(&a.maps[tid]);
        write_job_table(jobs, tid, JobSegment{a.prm.pi, mp, ps, 0, 0},
                        JobSegment{a.prm.q_target + net * q_stride, mt, qsh, net, 0},
                        JobSegment{online ? a.prm.q + net * q_stride : nullptr, mq, qsh, net, 0},
                        JobSegment{nullptr, nullptr, qsh, 0, 0});
    }
    float *head_pi = sm + pl.off_heads, *head_qt = head_pi + head_floats(ps.hidden, 2 * A),
          *head_q = head_qt + head_floats(qsh.hidden, 1);
    stage_head(head_qt, qsh, a.prm.q_target + net * q_stride);
    if (online) stage_head(head_q, qsh, a.prm.q + net * q_stride);
    WeightPipe pipe;
    // ---- policy over the P rows
    {
        const int S4 = round_up(S, 4);
        const int RPp = round_up(RPt, PASS_ROWS);
        // Long windows (get_l_probs over a burn-in, sac_base.py:2563-2568): the policy rows are the bulk of
        // the pass, so the E cluster ranks each run a slice of them and hand their head outputs to the
        // others over distributed shared memory instead of all running every row.
        const int NP = RPp / PASS_ROWS;
        const bool share = E > 1 && NP >= 2 * E;
        const int p_lo = share ? (NP * net / E) * PASS_ROWS : 0;
        const int p_hi = share ? (NP * (net + 1) / E) * PASS_ROWS : RPp;
        for (int i = tid + p_lo * S4; i < p_hi * S4; i += NT) {
            const int r = i / S4, col = i - r * S4;
            float v = 0.f;
            if (r < RP && col < S) {
                const int e = r / Lp, tt = r - e * Lp;
                v = st_p[((int64_t)(e0 + e) * L + t0 + tt) * S + col];
            } else if (r < RPt && col < S) {
                const int rv = r - RP, e = rv / (n + 1), k = rv - e * (n + 1);
                v = st_v[((int64_t)(e0 + e) * L + b + k) * S + col];
            }
            xin[r * lda + col] = v;
        }
        if (late_wait) pdl_wait();  // the policy's Adam step is complete and flushed from here on
        stage_head(head_pi, ps, a.prm.pi);
        __syncthreads();
        pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, pl.n_jobs);
        ASAC_PHASE(0, 1);
        // pi_handoff (host: train pass, one unshared policy pass): rank 0 leaves every layer's pre-activations of the
        // pass in the tile's slice of grad_pi_part for the policy backward
        const LayerBufs z_out = (!post && a.pi_handoff && net == 0)
                                    ? LayerBufs(a.wrk.grad_pi_part + (int64_t)blockIdx.x * net_stride(ps), PASS_ROWS * lda)
                                    : LayerBufs(nullptr);
        float *h = net_trunk_forward(ps, pipe, xin + p_lo * lda, bufA + p_lo * lda, bufB + p_lo * lda, nullptr, z_out,
                                     lda, p_hi - p_lo, part);
        head_forward(h, lda, ps.hidden, head_pi, head_pi + 2 * A * ps.hidden, 2 * A, min(RPt, p_hi) - p_lo,
                     ho + p_lo * 2 * A);
        __syncthreads();
        cluster.barrier_wait();  // (arrived at kernel entry: long complete) every rank of the cluster is running
        if (share) {
            cluster.sync();
            const int lo = p_lo * 2 * A, hi = min(RPt, p_hi) * 2 * A;
            for (int q = 0; q < E; ++q) {
                if (q == net) continue;
                float *remote = cluster.map_shared_rank(ho, q);
                for (int i = lo + tid; i < hi; i += NT) remote[i] = ho[i];
            }
            cluster.sync();
        }
    }

    ASAC_PHASE(0, 2);
    // ---- per P row: distribution, sampled action, log-probs, IS ratio, pi_probs, alpha terms
    // Stage A, one thread per (row, action dim): policy.py:169.  Stage B, three independent chains per
    // row on three different WARPS (value-row log-prob / IS ratio + pi_probs / alpha terms): as one
    // thread per row this block was 4.6 us of serial transcendentals with 8 of 512 threads busy.
    float alpha_term = 0.f, alpha_loss = 0.f;
    const float log_alpha = a.prm.log_alpha[0];
    for (int i = tid; i < RPt * A; i += NT) {
        const int r = i / A, j = i - r * A;
        float *hr = ho + r * 2 * A;
        const float m = hr[j], s = hr[A + j];
        hr[j] = policy_loc(m);
        hr[A + j] = policy_scale(s);
    }
    __syncthreads();
    const int RP32 = round_up(RP, 32);
    for (int idx = tid; idx < 3 * RP32; idx += NT) {
        const int part = idx / RP32, r = idx - part * RP32;
        if (r >= RP) continue;
        const int e = r / Lp, tt = r - e * Lp, t = t0 + tt, eg = e0 + e;
        const float *hr = ho + r * 2 * A;
        // the policy's output on the value row of (e, t): the same row, or the extra row on st_v
        const float *hv = (split && t >= b) ? ho + (RP + e * (n + 1) + (t - b)) * 2 * A : hr;
        if (part == 0 && t >= b) {  // value row k = t - b
            const int k = t - b, rv = e * (n + 1) + k;
            hr = hv;
            const float *eps = (post ? a.bat.eps_td : a.bat.eps_y) + ((int64_t)eg * (n + 1) + k) * A;
            float corr = 0.f;
            for (int j = 0; j < A; ++j) {
                const float x = hr[j] + eps[j] * hr[A + j];  // Normal.rsample: loc + eps * scale
                xs[rv * A + j] = x;
                corr += logf(squash_floor(x));
            }
            float lp_sum = 0.f;
            for (int j = 0; j < A; ++j) {
                float lp = normal_log_prob(xs[rv * A + j], hr[j], hr[A + j]) - corr;  // operators.py:12-14
                if (lp == INFINITY) lp = 0.f;                                         // operators.py:23
                lp_sum += lp;
            }
            logp[rv] = lp_sum;
        }
        if (part == 1 && use_is && t < L - 1) {  // pi / mu of the stored action (sac_base.py:1450-1455, 1159-1189)
            const float *act = a.bat.actions + ((int64_t)eg * c.bn_stride + t) * A;
            float fl = 1.f;
            for (int j = 0; j < A; ++j) fl *= squash_floor(atanhf(fminf(fmaxf(act[j], -0.999f), 0.999f)));
            float pi_prod = 1.f, mu_prod = 1.f;
            for (int j = 0; j < A; ++j) {
                const float xa = atanhf(fminf(fmaxf(act[j], -0.999f), 0.999f));
                float pj = expf(normal_log_prob(xa, hr[j], hr[A + j])) / fl;  // operators.py:17-19
                float mj;
                if (post) {
                    if (net == 0) a.wrk.pi_probs[((int64_t)eg * (L - 1) + t) * A + j] = pj;
                    mj = pj;
                    // _get_y inside _get_td_error: pi from the target states' policy, mu := pi_probs
                    if (split && t >= b) pj = expf(normal_log_prob(xa, hv[j], hv[A + j])) / fl;
                } else {
                    mj = a.bat.mu_probs[((int64_t)eg * c.bn_stride + t) * A + j];
                }
                if (isinf(pj)) pj = 1.f;  // prod_prob, operators.py:27-31
                if (isinf(mj)) mj = 1.f;
                pi_prod *= pj;
                mu_prod *= mj;
            }
            if (isinf(pi_prod) || isnan(pi_prod)) pi_prod = 1.f;
            if (isinf(mu_prod) || isnan(mu_prod)) mu_prod = 1.f;
            if (t >= b) ratio[e * n + (t - b)] = pi_prod / fmaxf(mu_prod, 1e-8f);  // sac_base.py:1275
        }
        if (part == 2 && post && c.use_auto_alpha && t == b) {  // sac_base.py:1931-1939
            const float *eps = a.bat.eps_alpha + (int64_t)eg * A;
            float corr = 0.f;
            for (int j = 0; j < A; ++j) corr += logf(squash_floor(hr[j] + eps[j] * hr[A + j]));
            float lp_sum = 0.f;
            int valid = 0;
            for (int j = 0; j < A; ++j) {
                const float x = hr[j] + eps[j] * hr[A + j];  // Normal.sample == torch.normal(loc, scale)
                float lp = normal_log_prob(x, hr[j], hr[A + j]) - corr;
                if (lp != INFINITY) ++valid; else lp = 0.f;
                lp_sum += lp;
            }
            const float target = c.target_c_alpha * (float)(-valid);
            const float term = -lp_sum - target;
            alpha_term += term;
            alpha_loss += log_alpha * term;
        }
    }
    __syncthreads();

    ASAC_PHASE(0, 3);
    // ---- critic inputs: V rows [state(e, b+k), tanh(x)], then S rows [state(e, b), stored action]
    const int K0 = S + A, K04 = round_up(K0, 4);
    const int RVp = round_up(RV, PASS_ROWS);
    const int s_row0 = post ? RVp : RV;  // post: S rows go through the online critics separately
    const int rows_all = round_up(s_row0 + RS, PASS_ROWS);
    for (int i = tid; i < rows_all * K04; i += NT) {
        const int r = i / K04, col = i - r * K04;
        float v = 0.f;
        if (r < RV) {
            const int e = r / (n + 1), k = r - e * (n + 1);
            if (col < S) v = st_v[((int64_t)(e0 + e) * L + b + k) * S + col];
            else if (col < K0) v = tanhf(xs[r * A + (col - S)]);
        } else if (r >= s_row0 && r < s_row0 + RS) {
            const int e = r - s_row0;
            if (col < S) v = st_p[((int64_t)(e0 + e) * L + b) * S + col];
            else if (col < K0) v = a.bat.actions[((int64_t)(e0 + e) * c.bn_stride + b) * A + (col - S)];
        }
        xin[r * lda + col] = v;
    }
    __syncthreads();

    ASAC_PHASE(0, 4);
    // ---- target critic `net` over the V rows (+ S rows for the clipped loss)
    {
        const int rq = need_tq ? RV + RS : RV;
        const int rqp = round_up(rq, PASS_ROWS);
        float *h = net_trunk_forward(qsh, pipe, xin, bufA, bufB, nullptr, nullptr, lda, rqp, part);
        float *qo = (h == bufA ? bufB : bufA);  // free buffer: head outputs [rq]
        head_forward(h, lda, qsh.hidden, head_qt, head_qt + qsh.hidden, 1, rq, qo);
        __syncthreads();
        for (int r = tid; r < rq; r += NT) {
            if (r < RV) qmin[r] = qo[r];
            else a.wrk.tq[(int64_t)net * B + e0 + (r - RV)] = qo[r];
        }
        __syncthreads();
    }
    ASAC_PHASE(0, 5);
    // ---- post: online critic `net` over the S rows (sac_base.py:2211-2216) unless the policy backward left them
    if (online) {
        float *h = net_trunk_forward(qsh, pipe, xin + s_row0 * lda, bufA, bufB, nullptr, nullptr, lda,
                                     round_up(RS, PASS_ROWS), part);
        head_forward(h, lda, qsh.hidden, head_q, head_q + qsh.hidden, 1, RS, qs);
        __syncthreads();
    }
    ASAC_PHASE(0, 6);
    // ---- ensemble combine on rank 0 over distributed shared memory, in member order
    float *qmin2 = ensemble_subset(c) && a.bat.ensemble_perms ? sm + pl.off_qmin2 : qmin;
    if (qmin2 == qmin && !online && a.push_combine) {
        // all members, nothing else to fetch: ranks 1.. PUSH their value rows into rank 0's shared memory and leave
        // after ONE cluster barrier; rank 0 takes the minimum from its own memory (a pull costs a second barrier,
        // to keep the remote memory alive, and a remote-load round trip: 1.3 us per pass)
        const int stride = round_up(pl.rows_max, 4);
        if (net != 0) {
            float *dst = cluster.map_shared_rank(sm + pl.off_qpush, 0) + (net - 1) * stride;
            for (int r = tid; r < RV; r += NT) dst[r] = qmin[r];
        }
        cluster.sync();
        if (net != 0) return;
        for (int i = 1; i < E; ++i) {
            const float *src = sm + pl.off_qpush + (i - 1) * stride;
            for (int r = tid; r < RV; r += NT) qmin[r] = fminf(qmin[r], src[r]);  // sac_base.py:1439-1442, member order
        }
        if (post)
            for (int i = tid; i < E * TBa; i += NT) {
                const int m = i / TBa, e = i - m * TBa;
                qs[m * TB + e] = __ldcg(a.wrk.tq + (int64_t)m * B + e0 + e);
            }
        __syncthreads();
    } else {
    cluster.sync();
    if (net == 0) {
        combine_value_rows(cluster, c, a.bat.ensemble_perms, post ? 3 : 0, qmin, qmin2, RV);
        if (online)
            for (int i = 1; i < E; ++i) {
                const float *rqs = cluster.map_shared_rank(qs, i);
                for (int e = tid; e < TBa; e += NT) qs[i * TB + e] = rqs[e];
            }
        else if (post)
            for (int i = tid; i < E * TBa; i += NT) {
                const int m = i / TBa, e = i - m * TBa;
                qs[m * TB + e] = __ldcg(a.wrk.tq + (int64_t)m * B + e0 + e);
            }
    }
    cluster.sync();  // remote shared memory stays alive until rank 0 has read it
    if (net != 0) return;
    }

    ASAC_PHASE(0, 7);
    // ---- per batch element: V, v-trace, y (sac_base.py:1244-1295, 1444-1464)
    // y is linear in alpha: V_k = qmin_k - alpha * logp_k.  The train pass knows alpha and writes y;
    // the post pass runs BEFORE the alpha Adam step of the same train() call but _get_td_error uses
    // the UPDATED alpha (sac_base.py:2115-2116 precede :2571), so it writes the two functionals
    // (critic part with rewards, log-prob part) and k_alpha_td combines them after the alpha step.
    if (tid < TBa) {
        const int e = tid, eg = e0 + e;
        const float alpha = expf(log_alpha);
        float q_prev = qmin[e * (n + 1)], l_prev = logp[e * (n + 1)];
        float v_prev = q_prev - alpha * l_prev;
        const float v0 = v_prev, q0 = q_prev, l0 = l_prev;
        float sum = 0.f, sum_q = 0.f, sum_l = 0.f, cprod = 1.f;
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const float q_next = qmin2[e * (n + 1) + k + 1], l_next = logp[e * (n + 1) + k + 1];  // "next rows" subset
            const float v_next = q_next - alpha * l_next;
            const int64_t idx = (int64_t)eg * c.bn_stride + b + k;
            const float nd = a.bat.dones[idx] ? 0.f : 1.f;
            const float rew = a.bat.rewards[idx];
            float td = rew + (c.gamma * nd) * v_next - v_prev;
            float td_q = rew + (c.gamma * nd) * q_next - q_prev;
            float td_l = (c.gamma * nd) * l_next - l_prev;
            td = c.gamma_ratio[k] * td;
            td_q = c.gamma_ratio[k] * td_q;
            td_l = c.gamma_ratio[k] * td_l;
            if (use_is) {
                td = c.lambda_ratio[k] * td;
                td_q = c.lambda_ratio[k] * td_q;
                td_l = c.lambda_ratio[k] * td_l;
                const float is = ratio[e * n + k];
                const float rho = fminf(is, c.v_rho);
                td = (cprod * rho) * td;
                td_q = (cprod * rho) * td_q;
                td_l = (cprod * rho) * td_l;
                cprod = cprod * fminf(is, c.v_c);
            }
            const float keep = (a.bat.last_masks[idx] | a.bat.padding_masks[idx]) ? 0.f : 1.f;
            sum += td * keep;
            sum_q += td_q * keep;
            sum_l += td_l * keep;
            q_prev = qmin[e * (n + 1) + k + 1];  // the same row as V_k of the next term: the "current rows" subset
            l_prev = l_next;
            v_prev = qmin2 == qmin ? v_next : q_prev - alpha * l_prev;
        }
        if (!post) {
            a.wrk.y[eg] = v0 + sum;
        } else {
            float *parts = a.wrk.post_parts + (int64_t)eg * (2 + E);
            parts[0] = q0 + sum_q;
            parts[1] = l0 + sum_l;
            for (int i = 0; i < E; ++i) parts[2 + i] = qs[i * TB + e];
        }
    }
    if (post && c.use_auto_alpha) {
        const float s0 = block_sum(alpha_term, red);
        const float s1 = block_sum(alpha_loss, red);
        if (tid == 0) {
            a.wrk.grad_alpha_part[blockIdx.x * 2 + 0] = s0;
            a.wrk.grad_alpha_part[blockIdx.x * 2 + 1] = s1;
        }
    }
    ASAC_PHASE(0, 31);
}

#include "sac_tc.cuh"  // the same passes on the tcgen05 layer engine (hidden width 64)

// ------------------------------------------------------------------------------------ critic
// grid (n_tiles, E): forward of Q_i on (s_b, a_b), clipped double loss, backward
// (sac_base.py:1516, 1539-1570).  Partial gradients are sums over the tile's rows of
// d(sum_i mean_B loss_i)/d theta_i.
__global__ void __launch_bounds__(NT) k_q_backward(const __grid_constant__ SacArgs a) {
    // Programmatic dependent launch: the forward pass reads nothing the value pass writes (parameters, batch), so it
    // runs while the predecessor drains; griddepcontrol.wait sits in front of the first read of y / tq.  (With a
    // trained representation the states come from the GRU forward TWO kernels ahead: the value pass in between waits
    // for it before it triggers this launch, so they are complete and visible here too.)  Stand-alone launches
    // (asac_sac_q_backward after arbitrary work of the caller) wait first.
    if (!a.late_wait) pdl_wait();
    pdl_trigger();
    warm_kernel_params(a);
    ASAC_PHASE(1, 0);
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const AsacSacConfig &c = a.cfg;
    const int tid = threadIdx.x;
    const int B = c.batch, L = c.seq_len, b = c.burn_in, S = c.state_size, A = c.action_size, E = c.ensemble;
    const int TB = a.tile_batch, e0 = blockIdx.x * TB, TBa = min(TB, B - e0), net = blockIdx.y;
    const GradPlan &pl = *reinterpret_cast<const GradPlan *>(a.plan);
    const int lda = pl.lda, R = PASS_ROWS;
    const NetShape qsh = q_shape(c);
    const int d = qsh.depth, H = qsh.hidden;
    const int64_t q_stride = net_stride(qsh);
    const float *prm = a.prm.q + net * q_stride;
    float *gout = a.wrk.grad_q_part + ((int64_t)blockIdx.x * E + net) * q_stride;

    const LayerBufs px(sm + pl.off_px, R * lda), pz(sm + pl.off_pz, R * lda);
    float *g[3] = {sm + pl.off_g0, sm + pl.off_g1, sm + pl.off_g2};
    float *qout = sm + pl.off_small, *dq = qout + R, *red = sm + pl.off_red, *part = sm + pl.off_part;
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    if (tid < 32) {  // one lane per job
        const CUtensorMap *mq = a.use_tma ? &a.maps[0] : nullptr;
        if (a.use_tma && tid == 0) prefetch_tensormap(&a.maps[0]);
        write_job_table(jobs, tid, JobSegment{prm, mq, qsh, net, 0}, JobSegment{prm, mq, qsh, net, 1},
                        JobSegment{nullptr, nullptr, qsh, 0, 0}, JobSegment{nullptr, nullptr, qsh, 0, 0});
    }
    float *head_q = sm + pl.off_heads + head_floats(c.pi_hidden, 2 * A);
    ASAC_PHASE(1, 8);
    stage_head(head_q, qsh, prm);
    __syncthreads();
    ASAC_PHASE(1, 9);
    WeightPipe pipe;
    pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, pl.n_jobs);
    ASAC_PHASE(1, 10);

    const int K0 = S + A, K04 = round_up(K0, 4);
    for (int i = tid; i < R * K04; i += NT) {
        const int r = i / K04, col = i - r * K04;
        float v = 0.f;
        if (r < TBa) {
            if (col < S) v = a.bat.states[((int64_t)(e0 + r) * L + b) * S + col];
            else if (col < K0) v = a.bat.actions[((int64_t)(e0 + r) * c.bn_stride + b) * A + (col - S)];
        }
        px[0][r * lda + col] = v;
    }
    __syncthreads();
    ASAC_PHASE(1, 1);
    net_trunk_forward(qsh, pipe, px[0], nullptr, nullptr, px, pz, lda, R, part);
    ASAC_PHASE(1, 2);
    head_forward(px[d], lda, H, head_q, head_q + H, 1, TBa, qout);
    __syncthreads();

    pdl_wait();  // y, tq: the value pass is complete and flushed from here on
    float loss = 0.f;
    if (tid < R) {
        float gq = 0.f;
        if (tid < TBa) {
            const int eg = e0 + tid;
            const float q = qout[tid], y = a.wrk.y[eg];
            const float w = (c.use_priority && a.bat.priority_is) ? a.bat.priority_is[eg] : 1.f;
            float l, gl;
            if (c.clip_epsilon > 0.f) {
                const float tq = a.wrk.tq[(int64_t)net * B + eg];
                const float diff = q - tq;
                const float cl = fminf(fmaxf(diff, -c.clip_epsilon), c.clip_epsilon);
                const float cq = tq + cl;
                const float la = (cq - y) * (cq - y), lb = (q - y) * (q - y);
                const float ga = (diff >= -c.clip_epsilon && diff <= c.clip_epsilon) ? 2.f * (cq - y) : 0.f;
                const float gb = 2.f * (q - y);
                l = fmaxf(la, lb);
                gl = la > lb ? ga : (la < lb ? gb : 0.5f * (ga + gb));  // torch.maximum splits ties
            } else {
                l = (q - y) * (q - y);
                gl = 2.f * (q - y);
            }
            loss = l * w;
            gq = (gl * w) / (float)B;
            a.wrk.q_val[(int64_t)net * B + eg] = q;
        }
        dq[tid] = gq;
    }
    loss = block_sum(loss, red);  // contains the barrier that publishes dq
    if (tid == 0) a.wrk.loss_q[blockIdx.x * E + net] = loss;

    // head backward, then the ResBlocks in reverse
    ASAC_PHASE(1, 3);
    // (dZ of the last ResBlock comes out of the head backward, the layers below get theirs from the input-gradient epilogue)
    head_backward(dq, 1, px[d], lda, H, head_q, R, gout + net_w_off(qsh, d), gout + net_b_off(qsh, d), g[0], lda,
                  pz[d - 1], g[1], TBa);
    ASAC_PHASE(1, 4);
    int cur = 0;
    __syncthreads();
#pragma unroll 1
    for (int l = d - 1; l >= 0; --l) {
        const int K = net_k(qsh, l);
        float *dY = g[cur], *dZ = g[(cur + 1) % 3], *dX = g[(cur + 2) % 3];
        layer_weight_grad(dZ, lda, px[l], lda, H, K, TBa, gout + net_w_off(qsh, l), gout + net_b_off(qsh, l));
        if (l > 0) {
            const float *Ws, *bs;
            const bool swz = pipe_front_swizzled(pipe);
            pipe_acquire(pipe, Ws, bs);
            // dX -> the next dY; the next dZ = dX * gelu'(z_{l-1}) lands where this dY was (the buffer rotation's slot)
            layer_input_grad(H, dZ, lda, Ws, dY, dX, true, part, swz, pz[l - 1], dY, TBa);  // ends with a CTA barrier
            pipe_release(pipe);
            cur = (cur + 2) % 3;
        } else if (a.wrk.grad_state) {
            // trained representation: d loss_i / d state[:, b] (the state columns of the first layer's
            // input gradient; target_c_q sees state.detach(), sac_base.py:1541)
            const float *W0 = prm + net_w_off(qsh, 0);
            for (int t = tid; t < TBa * S; t += NT) {
                const int r = t / S, k = t - r * S;
                float s = 0.f;
                for (int h = 0; h < H; ++h) s = fmaf(dZ[r * lda + h], __ldg(W0 + (int64_t)h * K0 + k), s);
                if (K0 == H) s += dY[r * lda + k];  // residual first block
                a.wrk.grad_state[((int64_t)net * B + e0 + r) * S + k] = s;
            }
        }
    }
    ASAC_PHASE(1, 31);
    gt_stamp(0, false);
}

// ------------------------------------------------------------------------------------ policy
// grid (n_tiles, E), cluster (1, E, 1): policy forward on s_b, rsample, critic `rank` on
// (s_b, tanh x), min over the cluster, backward through the own critic to the action, sum of
// the action gradients on rank 0, which then runs the policy backward (sac_base.py:1882-1908).
__global__ void __launch_bounds__(NT) k_policy_backward(const __grid_constant__ SacArgs a) {
    // Programmatic dependent launch: the policy forward reads the policy's parameters, the states and the noise —
    // nothing the critics' Adam step writes — so it runs while that kernel drains; griddepcontrol.wait sits in front
    // of the first read of the critics' parameters (their head and their weight jobs).  (With a trained
    // representation the re-encoded states come from the kernel just ahead: only the job table, the policy's head and
    // its weight jobs go before the wait.)  Stand-alone launches wait first.
    if (!a.late_wait) pdl_wait();
    pdl_trigger();
    gt_stamp(4, true); gt_stamp(5, false);
    warm_kernel_params(a);
    ASAC_PHASE(2, 0);
    cg::cluster_group cluster = cg::this_cluster();
    const int net = (int)cluster.block_rank();
    cluster.barrier_arrive();  // matched by barrier_wait() in front of the first remote store: every rank is running
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const AsacSacConfig &c = a.cfg;
    const int tid = threadIdx.x;
    const int B = c.batch, L = c.seq_len, b = c.burn_in, S = c.state_size, A = c.action_size, E = c.ensemble;
    const int TB = a.tile_batch, e0 = blockIdx.x * TB, TBa = min(TB, B - e0);
    const GradPlan &pl = *reinterpret_cast<const GradPlan *>(a.plan);
    const int lda = pl.lda, R = PASS_ROWS;
    const NetShape ps = pi_shape(c), qsh = q_shape(c);
    const int dp = ps.depth, Hp = ps.hidden, dqn = qsh.depth, Hq = qsh.hidden;
    const int64_t q_stride = net_stride(qsh), pi_stride = net_stride(ps);
    float *gout = a.wrk.grad_pi_part + (int64_t)blockIdx.x * pi_stride;

    const LayerBufs px(sm + pl.off_px, R * lda), pz(sm + pl.off_pz, R * lda);
    float *qin = sm + pl.off_qin;
    float *g[3] = {sm + pl.off_g0, sm + pl.off_g1, sm + pl.off_g2};
    // small: ho[R][2A] (m,s -> mu,sigma), xs[R][A], da[R][A], qv[E][R], dq[R], amin[R]
    float *ho = sm + pl.off_small, *xs = ho + R * 2 * A, *da = xs + R * A, *qv = da + R * A, *dq = qv + E * R;
    float *amin = dq + R, *dO = amin + R;  // dO aliases nothing: sized below
    float *red = sm + pl.off_red, *part = sm + pl.off_part;
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    const float *q_prm = a.prm.q + net * q_stride;
    // ranks != 0 leave before the policy backward: they must not have its weights in flight at exit
    const int n_jobs = pl.n_jobs - (net == 0 ? 0 : dp - 1);
    if (tid < 32) {  // one lane per job
        const CUtensorMap *mq = a.use_tma ? &a.maps[0] : nullptr, *mp = a.use_tma ? &a.maps[2] : nullptr;
        if (a.use_tma && (tid == 0 || tid == 2)) prefetch_tensormap(&a.maps[tid]);
        write_job_table(jobs, tid, JobSegment{a.pi_handoff ? nullptr : a.prm.pi, mp, ps, 0, 0},
                        JobSegment{q_prm, mq, qsh, net, 0}, JobSegment{q_prm, mq, qsh, net, 1},
                        JobSegment{net == 0 ? a.prm.pi : nullptr, mp, ps, 0, 1});
    }
    float *head_pi = sm + pl.off_heads, *head_q = head_pi + head_floats(Hp, 2 * A);
    stage_head(head_pi, ps, a.prm.pi);
    __syncthreads();
    WeightPipe pipe;
    pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, n_jobs,
              a.pi_handoff ? 0 : dp);  // before the wait: the policy trunk only (nothing with the hand-off)

    ASAC_PHASE(2, 1);
    // ---- policy forward (saved); with a trained representation: on the re-encoded states (sac_base.py:2107-2113)
    const float *st = a.bat.states_post ? a.bat.states_post : a.bat.states;
    const int S4 = round_up(S, 4);
    if (a.cfg.rep_kind != 0) pdl_wait();  // states_post: the GRU forward just ahead is complete from here on
    for (int i = tid; i < R * S4; i += NT) {
        const int r = i / S4, col = i - r * S4;
        px[0][r * lda + col] = (r < TBa && col < S) ? st[((int64_t)(e0 + r) * L + b) * S + col] : 0.f;
    }
    __syncthreads();
    if (a.pi_handoff) {
        // the saved forward from the train value pass's pre-activations (row e * (L - b) of its pass is s_b of batch
        // element e): y = gelu(z) (+ x), the expression of the layer routine, element by element — same bits
        const float *zg = a.wrk.grad_pi_part + (int64_t)blockIdx.x * pi_stride;
        const int Lv = L - b;
        const bool res0 = net_k(ps, 0) == Hp;
        for (int i = tid; i < R * Hp; i += NT) {
            const int r = i / Hp, j = i - r * Hp;
            float x = res0 ? px[0][r * lda + j] : 0.f;
#pragma unroll 1
            for (int l = 0; l < dp; ++l) {
                const float z = r < TBa ? __ldcg(zg + ((int64_t)l * R + r * Lv) * lda + j) : 0.f;
                float y = gelu_erf(z);
                if (l > 0 || res0) y = y + x;
                pz[l][r * lda + j] = z;
                px[l + 1][r * lda + j] = y;
                x = y;
            }
        }
        __syncthreads();
    } else {
        net_trunk_forward(ps, pipe, px[0], nullptr, nullptr, px, pz, lda, R, part);
    }
    head_forward(px[dp], lda, Hp, head_pi, head_pi + 2 * A * Hp, 2 * A, TBa, ho);
    __syncthreads();

    pdl_wait();  // the critics' Adam step is complete and flushed from here on
    gt_stamp(6, false);
    pipe_open_all(pipe);  // (the weight copies are in flight while the head is loaded)
    stage_head(head_q, qsh, q_prm);
    ASAC_PHASE(2, 2);
    // ---- sample, critic input
    const int K0 = S + A, K04 = round_up(K0, 4);
    // rows [TBa, 2 TBa) of the same pass: (s_b, stored action) for the td error of the post pass (q_sb_handoff)
    const bool handoff = a.q_sb_handoff && 2 * TBa <= R;
    for (int i = tid; i < R * K04; i += NT) {
        const int r = i / K04, col = i - r * K04;
        float v = 0.f;
        if (r < TBa) {
            if (col < S) {
                v = px[0][r * lda + col];
            } else if (col < K0) {
                const int j = col - S;
                const float mu = policy_loc(ho[r * 2 * A + j]), sg = policy_scale(ho[r * 2 * A + A + j]);
                const float x = mu + a.bat.eps_pi[(int64_t)(e0 + r) * A + j] * sg;
                xs[r * A + j] = x;
                v = tanhf(x);
            }
        } else if (handoff && r < 2 * TBa) {
            const int e = r - TBa;
            if (col < S) v = px[0][e * lda + col];
            else if (col < K0) v = a.bat.actions[((int64_t)(e0 + e) * c.bn_stride + b) * A + (col - S)];
        }
        qin[r * lda + col] = v;
    }
    for (int i = tid; i < R * A; i += NT) da[i] = 0.f;
    __syncthreads();

    ASAC_PHASE(2, 3);
    // ---- own critic forward (z saved), then the other members' values over DSMEM
    {
        const LayerBufs qz(sm + pl.off_qz, R * lda);
        float *h = net_trunk_forward(qsh, pipe, qin, g[0], g[1], nullptr, qz, lda, R, part);
        head_forward(h, lda, Hq, head_q, head_q + Hq, 1, handoff ? 2 * TBa : TBa, qv + net * R);
        __syncthreads();
        if (handoff && tid < TBa) a.wrk.tq[(int64_t)net * B + e0 + tid] = qv[net * R + TBa + tid];
    }
    // every rank PUSHES its member's values into the others' qv (remote stores, then one barrier: no remote-load
    // round trip behind it)
    cluster.barrier_wait();  // (arrived at kernel entry: long complete)
    for (int i = 0; i < E; ++i) {
        if (i == net) continue;
        float *rv = cluster.map_shared_rank(qv, i);
        if (tid < TBa) rv[net * R + tid] = qv[net * R + tid];
    }
    cluster.sync();
    if (tid < R) {
        int best = 0;
        if (tid < TBa) {
            if (ensemble_subset(c) && a.bat.ensemble_perms) {  // torch.min over stack[randperm[:Es]] (sac_base.py:1887-1894)
                const int32_t *perm = a.bat.ensemble_perms + 2 * E;
                best = perm[0];
                float m = qv[best * R + tid];
                for (int j = 1; j < c.ensemble_sample; ++j)
                    if (qv[perm[j] * R + tid] < m) { m = qv[perm[j] * R + tid]; best = perm[j]; }
            } else {
                float m = qv[tid];
                for (int i = 1; i < E; ++i)
                    if (qv[i * R + tid] < m) { m = qv[i * R + tid]; best = i; }
            }
        }
        amin[tid] = (float)best;
    }
    __syncthreads();

    ASAC_PHASE(2, 4);
    // ---- backward through the own critic to its action input; d loss / d q_min = -1/B
    {
        const int i = net;
        const float *prm = a.prm.q + i * q_stride;
        if (tid < R) dq[tid] = (tid < TBa && (int)amin[tid] == i) ? -1.f / (float)B : 0.f;
        __syncthreads();
        head_backward(dq, 1, nullptr, lda, Hq, head_q, R, nullptr, nullptr, g[0], lda,
                      sm + pl.off_qz + (dqn - 1) * R * lda, g[1], TBa);
        int cur = 0;
        __syncthreads();
#pragma unroll 1
        for (int l = dqn - 1; l >= 0; --l) {
            float *dY = g[cur], *dZ = g[(cur + 1) % 3], *dX = g[(cur + 2) % 3];
            if (l > 0) {
                const float *Ws, *bs;
                const bool swz = pipe_front_swizzled(pipe);
                pipe_acquire(pipe, Ws, bs);
                layer_input_grad(Hq, dZ, lda, Ws, dY, dX, true, part, swz, sm + pl.off_qz + (l - 1) * R * lda, dY,
                                 TBa);  // ends with a CTA barrier
                pipe_release(pipe);
                cur = (cur + 2) % 3;
            } else {
                // first layer: only the action columns of the input gradient are needed
                const float *W0 = prm + net_w_off(qsh, 0);
                for (int t = tid; t < TBa * A; t += NT) {
                    const int r = t / A, j = t - r * A;
                    float s = 0.f;
                    for (int h = 0; h < Hq; ++h) s = fmaf(dZ[r * lda + h], __ldg(W0 + (int64_t)h * K0 + S + j), s);
                    if (K0 == Hq) s += dY[r * lda + S + j];  // residual first block
                    da[r * A + j] += s;
                }
            }
        }
        __syncthreads();
    }
    ASAC_PHASE(2, 5);
    // ---- sum of the members' action gradients on rank 0, in member order: ranks 1.. push theirs and leave
    float *da_push = dO + R * 2 * A + 3 * R * A;  // behind dO and the three loss-term arrays in the `small` region
    if (net != 0) {
        float *dst = cluster.map_shared_rank(da_push, 0) + (net - 1) * R * A;
        for (int t = tid; t < TBa * A; t += NT) dst[t] = da[t];
    }
    cluster.sync();
    if (net != 0) return;
    for (int i = 1; i < E; ++i)
        for (int t = tid; t < TBa * A; t += NT) da[t] += da_push[(i - 1) * R * A + t];

    ASAC_PHASE(2, 6);
    // ---- d loss / d (mean, logstd) pre-activations; loss and entropy sums
    // One thread per (row, action dimension) for everything that does not need the row's summed Jacobian term (the
    // gradients of the head's outputs among them), then one thread per row for the two ordered sums: the same
    // expressions in the same order as a single thread per row would evaluate them, with the ten transcendental
    // calls of each dimension running side by side instead of in a row (3.4 -> 1.9 us at A = 2).
    float loss = 0.f, ent = 0.f;
    const float alpha = expf(a.prm.log_alpha[0]);
    float *lfs = dO + R * 2 * A, *lpn = lfs + R * A, *e1s = lpn + R * A;  // behind dO in the plan's `small` region
    for (int i = tid; i < R * A; i += NT) {
        const int r = i / A, j = i - r * A;
        if (r >= TBa) {
            dO[r * 2 * A + j] = 0.f;
            dO[r * 2 * A + A + j] = 0.f;
            continue;
        }
        const float m = ho[r * 2 * A + j], s = ho[r * 2 * A + A + j];
        const float mu = policy_loc(m), sg = policy_scale(s);
        const float x = xs[r * A + j], eps = a.bat.eps_pi[(int64_t)(e0 + r) * A + j];
        lfs[i] = logf(squash_floor(x));
        lpn[i] = normal_log_prob(x, mu, sg);
        e1s[i] = 0.5f + LOG_SQRT_2PI + logf(sg);  // Normal.entropy
        // d/dx of alpha * (-A * log max(1 - tanh^2 x, 1e-2)): every action dim carries the summed
        // Jacobian term (operators.py:12-14), hence the factor A
        const float t = tanhf(x), one_m = 1.f - t * t;
        const float dcorr = one_m > 1e-2f ? 2.f * t : (one_m == 1e-2f ? t : 0.f);
        const float dx = (alpha * (float)A * dcorr) / (float)B + da[r * A + j] * one_m;
        // the Normal.log_prob(rsample) terms cancel analytically except -log(scale)
        const float dsig = dx * eps - alpha / ((float)B * sg);
        const float th = tanhf(m / 5.f);
        dO[r * 2 * A + j] = dx * (1.f - th * th);
        dO[r * 2 * A + A + j] = (s >= -20.f && s <= 0.5f) ? dsig * sg : 0.f;
    }
    __syncthreads();
    if (tid < TBa) {
        const int r = tid;
        float corr = 0.f;
        const float qm = qv[(int)amin[r] * R + r];
        for (int j = 0; j < A; ++j) corr += lfs[r * A + j];
        float lp_sum = 0.f;
        for (int j = 0; j < A; ++j) {
            float lp = lpn[r * A + j] - corr;
            if (lp == INFINITY) lp = 0.f;
            lp_sum += lp;
            const float e1 = e1s[r * A + j];
            ent += (e1 == INFINITY) ? 0.f : e1;
        }
        loss = alpha * lp_sum - qm;
    }
    loss = block_sum(loss, red);
    ent = block_sum(ent, red);
    if (tid == 0) {
        a.wrk.stats_pi[blockIdx.x * 2 + 0] = loss;
        a.wrk.stats_pi[blockIdx.x * 2 + 1] = ent;
    }

    ASAC_PHASE(2, 7);
    // ---- policy backward
    head_backward(dO, 2 * A, px[dp], lda, Hp, head_pi, R, gout + net_w_off(ps, dp), gout + net_b_off(ps, dp), g[0],
                  lda, pz[dp - 1], g[1], TBa);
    int cur = 0;
    __syncthreads();
#pragma unroll 1
    for (int l = dp - 1; l >= 0; --l) {
        const int K = net_k(ps, l);
        float *dY = g[cur], *dZ = g[(cur + 1) % 3], *dX = g[(cur + 2) % 3];
        layer_weight_grad(dZ, lda, px[l], lda, Hp, K, TBa, gout + net_w_off(ps, l), gout + net_b_off(ps, l));
        if (l > 0) {
            const float *Ws, *bs;
            const bool swz = pipe_front_swizzled(pipe);
            pipe_acquire(pipe, Ws, bs);
            layer_input_grad(Hp, dZ, lda, Ws, dY, dX, true, part, swz, pz[l - 1], dY, TBa);  // ends with a CTA barrier
            pipe_release(pipe);
            cur = (cur + 2) % 3;
        }
    }
    ASAC_PHASE(2, 31);
}

// ------------------------------------------------------------------------------------ optimiser
// ---- gradient exchange over NVLink peer memory (data-parallel learner)
// Every rank owns a symmetric receive buffer of 8-byte words  recv[parity 2][source rank][total]
// that all peers have mapped.  A thread that has reduced one element of the local gradient PUSHES
// the pair {epoch, value} as ONE 64-bit store into every rank's buffer (its own included) and
// then polls the words of all sources for the same element until they carry the step's epoch —
// the flag travels with the data (the "LL" idea of NCCL's low-latency protocol), so no fence, no
// separate flag round trip and no barrier sits between the backward pass and Adam: the cost is
// one NVLink write latency.  Sources are summed in rank order, so every rank applies
// bit-identical gradients.  Epochs only grow (optimizer step + 1) and the buffer alternates with
// the epoch's parity: a source can be at most one epoch ahead of a consumer, never two.
// (A first version used per-CTA flags behind __threadfence_system() + st.release.sys: two
// system-scope fences with NVLink writes in flight per CTA cost more than the NCCL all-reduce.)
struct PeerExchange {
    int world, rank;          // world <= 1: no exchange
    unsigned long long *recv[ASAC_MAX_PEERS];
    int64_t total;            // words per (parity, source) block
    int64_t off;              // where this gradient kind starts inside a block
};
__device__ __forceinline__ unsigned long long *peer_slot(const PeerExchange &x, int dst_rank, unsigned epoch,
                                                         int src_rank) {
    return x.recv[dst_rank] + ((int64_t)(epoch & 1u) * x.world + src_rank) * x.total + x.off;
}
__device__ __forceinline__ void peer_push(const PeerExchange &x, unsigned epoch, int64_t p, float v) {
    const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
    for (int q = 0; q < x.world; ++q) {
        unsigned long long *dst = peer_slot(x, q, epoch, x.rank) + p;
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
    }
}
// Polls are bounded: a peer that crashed, or ranks that disagree on the number of train() calls, must not
// hang the GPU inside a kernel.  After PEER_TIMEOUT_NS without the expected epoch the wait is abandoned,
// g_peer_timeouts is bumped (the host reads it through asac_peer_timeouts and raises) and the missing
// contribution counts as zero.
__device__ unsigned int g_peer_timeouts;
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// All sources are polled TOGETHER: the loads of a round are independent and issued back to back, so a round costs one
// L2 round trip whatever the world size (polling the sources one after the other — load, test, next — cost one
// round trip EACH: +0.6 us per rank and exchange, the growth of the step with N measured in rounds 1 and 2).  The
// sum runs over the sources in rank order, as before: bit-identical on every rank.
__device__ __forceinline__ float peer_sum(const PeerExchange &x, unsigned epoch, int64_t p) {
    const unsigned long long *src0 = peer_slot(x, x.rank, epoch, 0) + p;  // source q at src0 + q * x.total
    unsigned long long w[ASAC_MAX_PEERS];
    unsigned pending = x.world >= 32 ? 0xffffffffu : ((1u << x.world) - 1u);
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (true) {
#pragma unroll
        for (int q = 0; q < ASAC_MAX_PEERS; ++q)
            if ((pending >> q) & 1u)
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[q]) : "l"(src0 + (int64_t)q * x.total) : "memory");
#pragma unroll
        for (int q = 0; q < ASAC_MAX_PEERS; ++q)
            if (((pending >> q) & 1u) && (unsigned)(w[q] >> 32) == epoch) pending &= ~(1u << q);
        if (pending == 0) break;
        if (spins == 0) t0 = global_timer_ns();
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
            atomicAdd(&g_peer_timeouts, 1u);
#pragma unroll
            for (int q = 0; q < ASAC_MAX_PEERS; ++q)
                if ((pending >> q) & 1u) w[q] = (unsigned long long)epoch << 32;  // +0.0f
            break;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < ASAC_MAX_PEERS; ++q)
        if (q < x.world) s += __uint_as_float((unsigned)(w[q] & 0xFFFFFFFFull));
    return s;
}

struct AdamArgs {
    PeerExchange px;
    float *param, *m, *v;
    const float *part;      // partial gradients (or nullptr: read `grad`)
    float *grad;            // reduced gradient (written when part != nullptr and write_grad)
    const int64_t *step;    // optimizer step counter (value BEFORE this step)
    int64_t count;          // number of floats
    int64_t tile_stride;    // stride between tiles in `part`
    int n_tiles;
    int write_grad, do_adam;
    float grad_scale;       // applied to the reduced gradient (1/B folded elsewhere; 1/world for DDP)
    double lr;
};

constexpr int ADAM_PARAMS_PER_CTA = 64, ADAM_TILE_GROUPS = 4;

// Deterministic reduction of the per-tile partial gradients fused with torch.optim.Adam's
// single-tensor update (betas 0.9/0.999, eps 1e-8, no weight decay / amsgrad).  A CTA owns 64
// consecutive parameters; its 4 thread groups each sum a quarter of the tiles (coalesced 256-byte
// rows, loads of different tiles in flight together), the quarters are added in group order.
// The float64 bias corrections are computed once per CTA while the loads are in flight.
__global__ void __launch_bounds__(ADAM_PARAMS_PER_CTA *ADAM_TILE_GROUPS) k_reduce_adam(const AdamArgs a) {
    __shared__ float s_part[ADAM_TILE_GROUPS][ADAM_PARAMS_PER_CTA];
    __shared__ float s_bc[2];
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, pl = tid & (ADAM_PARAMS_PER_CTA - 1), tg = tid / ADAM_PARAMS_PER_CTA;
    const int64_t p = (int64_t)blockIdx.x * ADAM_PARAMS_PER_CTA + pl;
    float gr = 0.f;
    if (a.part) {
        if (p < a.count) {
            const int per = (a.n_tiles + ADAM_TILE_GROUPS - 1) / ADAM_TILE_GROUPS;
            const int t0 = tg * per, t1 = min(a.n_tiles, t0 + per);
#pragma unroll 8
            for (int t = t0; t < t1; ++t) gr += __ldcg(a.part + t * a.tile_stride + p);
        }
        s_part[tg][pl] = gr;
    }
    if (tid == ADAM_PARAMS_PER_CTA * ADAM_TILE_GROUPS - 1 && a.do_adam) {  // a thread of the last warp
        const double t = (double)(a.step[0] + 1);
        const double bc1 = 1.0 - beta_pow(ADAM_LN_BETA1, t), bc2 = 1.0 - beta_pow(ADAM_LN_BETA2, t);
        s_bc[0] = (float)(-(a.lr / bc1));
        s_bc[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    if (a.px.world > 1) {
        // local slice -> every rank's receive buffer, then flags, then the rank-ordered sum
        const unsigned epoch = (unsigned)(a.step[0] + 1);
        if (tg == 0 && p < a.count) {
            if (a.part) {
                gr = s_part[0][pl];
#pragma unroll
                for (int g = 1; g < ADAM_TILE_GROUPS; ++g) gr += s_part[g][pl];
            } else {
                gr = a.grad[p];
            }
            peer_push(a.px, epoch, p, gr);
        }
        if (tg != 0 || p >= a.count) return;
        gr = peer_sum(a.px, epoch, p);
        if (a.write_grad) a.grad[p] = gr;
    } else {
        if (tg != 0 || p >= a.count) return;
        if (a.part) {
            gr = s_part[0][pl];
#pragma unroll
            for (int g = 1; g < ADAM_TILE_GROUPS; ++g) gr += s_part[g][pl];
            if (a.write_grad) a.grad[p] = gr;
        } else {
            gr = a.grad[p];
        }
    }
    if (!a.do_adam) return;
    gr = gr * a.grad_scale;
    const float step_size = s_bc[0], bc2_sqrt = s_bc[1];
    const float w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
    float m = a.m[p], v = a.v[p];
    m = m + w1 * (gr - m);                      // exp_avg.lerp_(grad, 1 - beta1)
    v = v * 0.999f + (w2 * gr) * gr;            // mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + 1e-8f;
    a.m[p] = m;
    a.v[p] = v;
    a.param[p] = a.param[p] + (step_size * m) / denom;  // addcdiv_(exp_avg, denom, value=-step_size)
}

// alpha: a single scalar, gradient = (sum over tiles of the alpha terms) / B  (sac_base.py:1941-1948),
// then y' = yq - alpha_new * yl and td = mean_i |Q_i(s_b, a_b) - y'|  (sac_base.py:2223-2245).  One CTA.
// `staged` (shared memory, n_tiles floats) holds wrk.grad_alpha_part[t * 2], loaded by the whole CTA:
// one thread summing the tiles straight from global memory serialised n_tiles L2 round trips
// Executed by ONE thread.  With a peer exchange the scalar gradient travels like a 1-float slice.
// -> log_alpha after the call
__device__ __forceinline__ float alpha_reduce_adam(const AsacSacParams &prm, const AsacSacWork &wrk, int n_tiles,
                                                   int batch, int do_reduce, int do_adam, float grad_scale,
                                                   double lr, const float *staged, const PeerExchange *px = nullptr) {
    float log_alpha = do_adam ? prm.log_alpha[0] : 0.f;  // (a reduce-only launch carries no parameters)
    {
        float gr;
        if (do_reduce) {
            float s = 0.f;
            for (int t = 0; t < n_tiles; ++t) s += staged[t];
            gr = s / (float)batch;
            wrk.grad_alpha[0] = gr;
        } else {
            gr = wrk.grad_alpha[0];
        }
        if (px && px->world > 1) {
            const unsigned epoch = (unsigned)(prm.counters[3] + 1);
            peer_push(*px, epoch, 0, gr);
            gr = peer_sum(*px, epoch, 0);
            wrk.grad_alpha[0] = gr;
        }
        if (do_adam) {
            gr = gr * grad_scale;
            const double t = (double)(prm.counters[3] + 1);
            const double bc1 = 1.0 - beta_pow(ADAM_LN_BETA1, t), bc2 = 1.0 - beta_pow(ADAM_LN_BETA2, t);
            const float step_size = (float)(-(lr / bc1));
            const float bc2_sqrt = (float)sqrt(bc2);
            const float w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
            float m = prm.alpha_m[0], v = prm.alpha_v[0];
            m = m + w1 * (gr - m);
            v = v * 0.999f + (w2 * gr) * gr;
            const float denom = sqrtf(v) / bc2_sqrt + 1e-8f;
            prm.alpha_m[0] = m;
            prm.alpha_v[0] = v;
            log_alpha = log_alpha + (step_size * m) / denom;
            prm.log_alpha[0] = log_alpha;
        }
    }
    return log_alpha;
}

__global__ void __launch_bounds__(1024) k_alpha_td(const AsacSacParams prm, const AsacSacWork wrk, int n_tiles,
                                                   int batch, int ensemble, int do_reduce, int do_adam, int do_td,
                                                   float grad_scale, double lr) {
    __shared__ float s_alpha[1024];
    if (do_reduce) {
        for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) s_alpha[t] = __ldcg(wrk.grad_alpha_part + t * 2);
        __syncthreads();
    }
    if (threadIdx.x == 0 && (do_reduce || do_adam))
        alpha_reduce_adam(prm, wrk, n_tiles, batch, do_reduce, do_adam, grad_scale, lr, s_alpha);
    if (!do_td) return;
    __syncthreads();
    const float alpha = expf(__ldcg(prm.log_alpha));
    for (int e = threadIdx.x; e < batch; e += blockDim.x) {
        const float *parts = wrk.post_parts + (int64_t)e * (2 + ensemble);
        const float y = parts[0] - alpha * parts[1];
        float acc = 0.f;
        for (int i = 0; i < ensemble; ++i) acc += fabsf(parts[2 + i] - y);
        wrk.y_td[e] = y;
        wrk.td_error[e] = acc / (float)ensemble;
    }
}

// The tail of one train() in ONE CTA (batch <= 1024): alpha reduce + Adam (sac_base.py:1941-1948),
// y' and td error with the updated alpha (:2223-2245), PrioritizedReplayBuffer.update
// (replay_buffer.py:412-427) and the step / optimizer counters (sac_base.py:2607).
struct EpilogueArgs {
    PeerExchange px;
    float grad_scale;
    AsacSacParams prm;
    AsacSacWork wrk;
    int n_tiles, batch, ensemble, use_auto_alpha, counter_mask;
    double lr;
    float *nodes;
    int64_t capacity;
    int levels;
    const int64_t *store_ids, *data_ids;
    float td_min, td_max, per_alpha;
    double *per_state;
};
__global__ void __launch_bounds__(1024) k_step_epilogue(const __grid_constant__ EpilogueArgs a) {
    __shared__ TreeApplySmem s_apply;
    __shared__ float s_alpha[1024];
    __shared__ float s_log_alpha;
    pdl_wait();
    pdl_trigger();
    const int t = threadIdx.x;
    if (a.use_auto_alpha) {
        for (int i = t; i < a.n_tiles; i += blockDim.x) s_alpha[i] = __ldcg(a.wrk.grad_alpha_part + i * 2);
        __syncthreads();
        if (t == 0)  // (the updated value goes to the other threads through shared memory, not a global round trip)
            s_log_alpha = alpha_reduce_adam(a.prm, a.wrk, a.n_tiles, a.batch, 1, 1, a.grad_scale, a.lr, s_alpha, &a.px);
    } else if (t == 0) {
        s_log_alpha = __ldcg(a.prm.log_alpha);
    }
    __syncthreads();
    const float alpha = expf(s_log_alpha);
    bool active = t < a.batch;
    int slot = 0, bad = 0;
    float value = 0.f;
    if (active) {
        const float *parts = a.wrk.post_parts + (int64_t)t * (2 + a.ensemble);
        const float y = parts[0] - alpha * parts[1];
        float acc = 0.f;
        for (int i = 0; i < a.ensemble; ++i) acc += fabsf(parts[2 + i] - y);
        const float td = acc / (float)a.ensemble;
        a.wrk.y_td[t] = y;
        a.wrk.td_error[t] = td;
        const int64_t id = a.data_ids[t];
        slot = (int)(id & (a.capacity - 1));
        value = td_to_priority(td, a.td_min, a.td_max, a.per_alpha, &bad);
        active = (a.store_ids[slot] == id);
    }
    if (t < 8 && ((a.counter_mask >> t) & 1)) a.prm.counters[t] += 1;  // after the Adam step read counters[3]
    if (__syncthreads_or(bad)) {
        if (t == 0) a.per_state[3] = 1.0;  // the reference raises 'td_error has nan'
        return;
    }
    if (a.nodes) block_tree_apply(a.nodes, a.capacity, a.levels, slot, value, active, s_apply);  // else: deferred
}

// torch.randperm(E) x n_perms: Fisher-Yates, one thread per permutation
__global__ void k_ensemble_perms(int32_t *out, int n_perms, int E, uint64_t seed, const int64_t *counter) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_perms) return;
    int32_t v[ASAC_MAX_ENSEMBLE];
    for (int i = 0; i < E; ++i) v[i] = i;
    uint32_t r[4];
    for (int i = E - 1; i > 0; --i) {
        philox4(seed ^ 0xE5E3B1E5ull, (uint64_t)counter[0], ((uint64_t)p << 8) | (uint64_t)i, r);
        const int j = (int)(u01_double(r[0], r[1]) * (double)(i + 1));
        const int32_t t = v[i]; v[i] = v[j]; v[j] = t;
    }
    for (int i = 0; i < E; ++i) out[p * E + i] = v[i];
}

__global__ void k_bump(int64_t *counters, int mask) {
    if (threadIdx.x < 8 && ((mask >> threadIdx.x) & 1)) counters[threadIdx.x] += 1;
}

// sac_base.py:745-764: target = target * (1 - tau) + source * tau, gated on the global step
__global__ void __launch_bounds__(256) k_polyak(float *target, const float *source, int64_t count,
                                                const int64_t *counters, int per_step, float tau, float one_m,
                                                int force) {
    if (!force && (counters[0] % per_step) != 0) return;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    target[p] = target[p] * one_m + source[p] * tau;
}

__global__ void __launch_bounds__(256) k_fill_normal(float *out, int64_t n, uint64_t seed, const int64_t *counter,
                                                     int stream_id) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 4 outputs per thread
    if (q * 4 >= n) return;
    uint32_t r[4];
    philox4(seed ^ ((uint64_t)stream_id << 56), (uint64_t)counter[0], (uint64_t)q, r);
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)(r[2 * h] >> 8) + 1.f) * (1.f / 16777216.f);  // (0, 1]
        const float u2 = (float)(r[2 * h + 1] >> 8) * (1.f / 16777216.f);      // [0, 1)
        const float rad = sqrtf(-2.f * logf(u1));
        float sn, cs;
        sincospif(2.f * u2, &sn, &cs);
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
    }
    for (int i = 0; i < 4; ++i)
        if (q * 4 + i < n) out[q * 4 + i] = z[i];
}

struct PrefetchArgs {
    const char *base[8];
    int64_t lines[8];  // 128-byte lines per region (prefix sums in `first`)
    int64_t first[9];
    int n;
};
__global__ void __launch_bounds__(256) k_l2_prefetch(const PrefetchArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.first[a.n]) return;
    int r = 0;
    while (r + 1 < a.n && i >= a.first[r + 1]) ++r;
    const char *p = a.base[r] + (i - a.first[r]) * 128;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Actor side (sac_base.py:943-964, continuous branch of _choose_action): from the policy head's
// pre-activations to the squashed action and its per-dimension probability.
//   c_action = offline | tanh(mean) (disable_sample) | tanh(Normal(mu, sigma).sample())
//   prob_j   = exp(logN(x_j)) / prod_k max(1 - tanh(x_k)^2, 1e-2),  x = atanh(clamp(c_action, +-0.999))
// (utils/operators.py:17-19).  eps == nullptr: N(0,1) from Philox keyed by (seed, counter[0], row).
struct ActArgs {
    const float *pre;       // [rows, 2A]: mean, logstd pre-activations
    const float *eps;       // [rows, A] or null
    const float *offline;   // [rows, A] or null
    float *action, *prob;   // [rows, A]
    int64_t rows;
    int A, disable_sample;
    uint64_t seed;
    const int64_t *counter;
};
__global__ void __launch_bounds__(256) k_policy_sample(const ActArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    const int A = a.A;
    const float *pre = a.pre + r * 2 * A;
    float fl = 1.f;
    for (int j = 0; j < A; ++j) {
        const float mu = policy_loc(pre[j]), sg = policy_scale(pre[A + j]);
        float act;
        if (a.offline) {
            act = a.offline[r * A + j];
        } else if (a.disable_sample) {
            act = tanhf(mu);
        } else {
            float e;
            if (a.eps) {
                e = a.eps[r * A + j];
            } else {
                uint32_t rnd[4];
                philox4(a.seed ^ 0xAC7ull, (uint64_t)a.counter[0], (uint64_t)(r * A + j), rnd);
                const float u1 = ((float)(rnd[0] >> 8) + 1.f) * (1.f / 16777216.f);
                const float u2 = (float)(rnd[1] >> 8) * (1.f / 16777216.f);
                float sn, cs;
                sincospif(2.f * u2, &sn, &cs);
                e = sqrtf(-2.f * logf(u1)) * cs;
            }
            act = tanhf(mu + e * sg);  // torch.normal(mean, std)
        }
        a.action[r * A + j] = act;
        fl *= squash_floor(atanhf(fminf(fmaxf(act, -0.999f), 0.999f)));
    }
    for (int j = 0; j < A; ++j) {
        const float mu = policy_loc(pre[j]), sg = policy_scale(pre[A + j]);
        const float x = atanhf(fminf(fmaxf(a.action[r * A + j], -0.999f), 0.999f));
        a.prob[r * A + j] = expf(normal_log_prob(x, mu, sg)) / fl;
    }
}

// standalone stock-net forward (actor side / tests)
struct MlpArgs {
    const float *params, *x;
    float *out;
    NetShape s;
    int64_t rows;
    int rows_per_cta;
};
// shared-memory plan of k_mlp_forward (floats)
struct MlpPlan {
    int lda, wsz, off_a, off_b, off_ho, off_part, off_pipe, off_slots, n_slots, total;
};
__host__ __device__ __forceinline__ MlpPlan mlp_plan(const NetShape &s, int rows_per_cta) {
    MlpPlan p;
    p.lda = tile_lda(s.hidden, s.in_dim);
    p.wsz = round_up(tile_wsz(s.hidden, s.in_dim), 256);
    int o = rows_per_cta * p.lda;  // xin at 0
    p.off_a = o; o += rows_per_cta * p.lda;
    p.off_b = o; o += rows_per_cta * p.lda;
    p.off_ho = o; o += round_up(rows_per_cta * s.out_dim, 4);
    p.off_part = o; o += tile_part_floats(s.hidden);
    p.off_pipe = o; o += PIPE_HEADER_FLOATS;
    p.off_slots = o;
    o += 256;  // 1024-byte alignment slack of the slots
    p.n_slots = slots_that_fit(o, p.wsz, s.depth);
    p.total = o + p.n_slots * p.wsz;
    return p;
}
__global__ void __launch_bounds__(NT) k_mlp_forward(const MlpArgs a) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const int tid = threadIdx.x;
    const NetShape s = a.s;
    const int RC = a.rows_per_cta;
    const MlpPlan pl = mlp_plan(s, RC);
    const int lda = pl.lda;
    const int64_t r0 = (int64_t)blockIdx.x * RC;
    const int rows = (int)min((int64_t)RC, a.rows - r0);
    float *xin = sm, *bufA = sm + pl.off_a, *bufB = sm + pl.off_b, *ho = sm + pl.off_ho, *part = sm + pl.off_part;
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    if (tid == 0) push_trunk_jobs(jobs, 0, s, a.params);
    __syncthreads();
    WeightPipe pipe;
    pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, s.depth);
    const int K4 = round_up(s.in_dim, 4);
    const int rp = round_up(rows, PASS_ROWS);
    for (int i = tid; i < rp * K4; i += NT) {
        const int r = i / K4, col = i - r * K4;
        xin[r * lda + col] = (r < rows && col < s.in_dim) ? a.x[(r0 + r) * s.in_dim + col] : 0.f;
    }
    __syncthreads();
    float *h = net_trunk_forward(s, pipe, xin, bufA, bufB, nullptr, nullptr, lda, rp, part);
    head_forward(h, lda, s.hidden, a.params + net_w_off(s, s.depth), a.params + net_b_off(s, s.depth), s.out_dim, rows,
                 ho);
    __syncthreads();
    for (int i = tid; i < rows * s.out_dim; i += NT) a.out[r0 * s.out_dim + i] = ho[i];
}

#include "sac_discrete.cuh"  // discrete / hybrid action branches

}  // namespace asac

using namespace asac;

// ------------------------------------------------------------------------------------ host side
static int validate(const AsacSacConfig *c) {
    ASAC_REQUIRE(c != nullptr, "null config");
    ASAC_REQUIRE(c->batch > 0 && c->seq_len == c->burn_in + c->n_step + 1 && c->n_step >= 1 && c->burn_in >= 0,
                 "bad batch/sequence sizes (B=%d L=%d b=%d n=%d)", c->batch, c->seq_len, c->burn_in, c->n_step);
    ASAC_UNSUPPORTED(c->n_step > ASAC_MAX_NSTEP, "n_step %d > %d", c->n_step, ASAC_MAX_NSTEP);
    ASAC_UNSUPPORTED(c->ensemble < 1 || c->ensemble > ASAC_MAX_ENSEMBLE, "ensemble_q_num %d outside [1, %d]",
                     c->ensemble, ASAC_MAX_ENSEMBLE);
    ASAC_UNSUPPORTED(c->q_depth < 1 || c->q_depth > ASAC_MAX_DEPTH || c->pi_depth < 1 || c->pi_depth > ASAC_MAX_DEPTH,
                     "dense depth outside [1, %d]", ASAC_MAX_DEPTH);
    const int hs[2] = {c->q_hidden, c->pi_hidden};
    for (int h : hs)
        ASAC_UNSUPPORTED(!(h == 16 || h == 32 || h == 64 || h == 128), "hidden width %d not in {16, 32, 64, 128}", h);
    ASAC_REQUIRE(c->state_size > 0 && c->action_size > 0, "state/action size must be positive");
    ASAC_UNSUPPORTED(c->action_size > 64, "action_size %d > 64", c->action_size);
    ASAC_REQUIRE(c->bn_stride >= c->seq_len - 1, "bn_stride %d < L-1", c->bn_stride);
    ASAC_REQUIRE(c->update_target_per_step >= 1, "update_target_per_step < 1");
    return ASAC_OK;
}

static const int kSmemLimit = 227 * 1024;

extern "C" int asac_sac_tile_batch(const AsacSacConfig *c) {
    if (validate(c) != ASAC_OK) return ASAC_EINVAL;
    // The step is latency-bound at replay batch sizes: prefer >= 64 batch tiles (x E cluster ranks
    // >= 128 CTAs on the 148 SMs) over full 16-row tiles, then the largest tile that fits.
    int want = PASS_ROWS;
    while (want > 1 && c->batch / want < 64) want >>= 1;
    static const int forced = [] {  // experiments: ASAC_TILE_BATCH=8 / 16
        const char *e = getenv("ASAC_TILE_BATCH");
        return e ? atoi(e) : 0;
    }();
    if (forced > 0 && forced <= PASS_ROWS) want = forced;
    for (int tb = want; tb >= 1; tb >>= 1) {
        const int need0 = value_plan(*c, tb, 0).total * 4, need1 = value_plan(*c, tb, 1).total * 4;
        if (need0 <= kSmemLimit && need1 <= kSmemLimit) return tb;
    }
    set_error("asac_sac_tile_batch: sequence too long for the fused value pass");
    return ASAC_EUNSUPPORTED;
}

extern "C" int64_t asac_mlp_param_count(int in_dim, int hidden, int depth, int out_dim) {
    return net_count(NetShape{in_dim, hidden, depth, out_dim});
}
extern "C" int64_t asac_mlp_param_stride(int in_dim, int hidden, int depth, int out_dim) {
    return net_stride(NetShape{in_dim, hidden, depth, out_dim});
}

// ---- TMA tensor maps of the hidden -> hidden weights (cached per parameter buffer)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        const char *e = getenv("ASAC_TMA");
        if (e && e[0] == '0') return (TensorMapEncodeFn) nullptr;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (TensorMapEncodeFn)p;
    }();
    return fn;
}
struct CachedMap {
    const float *base;
    int hidden, depth, nets;
    int64_t stride;
    CUtensorMap map;
};
// dims {k = H, row = H, layer = depth - 1, member}: layer l >= 1 of member i starts at
// base + i * stride + net_w_off(shape, l); consecutive hidden -> hidden layers are H * (H + 1) floats apart
static bool weight_tensor_map(CUtensorMap *out, const float *base, const NetShape &s, int nets, int64_t stride) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc || s.depth < 2 || s.hidden < 32 || (((uintptr_t)base) & 15) != 0) return false;
    static std::mutex mu;
    static std::vector<CachedMap> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (const CachedMap &c : cache)
        if (c.base == base && c.hidden == s.hidden && c.depth == s.depth && c.nets == nets && c.stride == stride) {
            *out = c.map;
            return true;
        }
    const cuuint64_t H = (cuuint64_t)s.hidden;
    cuuint64_t dims[4] = {H, H, (cuuint64_t)(s.depth - 1), (cuuint64_t)nets};
    cuuint64_t strides[3] = {H * 4, H * (H + 1) * 4, (cuuint64_t)stride * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)s.hidden, 1, 1}, elem[4] = {1, 1, 1, 1};
    CachedMap c;
    c.base = base; c.hidden = s.hidden; c.depth = s.depth; c.nets = nets; c.stride = stride;
    const CUresult r = enc(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)(base + net_w_off(s, 1)), dims, strides, box,
                           elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() < 256) cache.push_back(c);
    *out = c.map;
    return true;
}

static int make_args(SacArgs &a, const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                     const AsacSacWork *wrk) {
    int rc = validate(cfg);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(prm && wrk, "null params/work");
    memset(&a.maps, 0, sizeof(a.maps));
    {
        const NetShape qs = q_shape(*cfg), ps = pi_shape(*cfg);
        const bool needq = qs.depth >= 2 && qs.hidden >= 32, needp = ps.depth >= 2 && ps.hidden >= 32;
        const bool okq = !needq || (weight_tensor_map(&a.maps[0], prm->q, qs, cfg->ensemble, net_stride(qs)) &&
                                    weight_tensor_map(&a.maps[1], prm->q_target, qs, cfg->ensemble, net_stride(qs)));
        const bool okp = !needp || weight_tensor_map(&a.maps[2], prm->pi, ps, 1, net_stride(ps));
        // one flag for the kernels: families without hidden -> hidden layers have no TMA jobs anyway
        // (tma_layer() in mlp_tile.cuh); a failed encode disables TMA for the whole launch
        a.use_tma = (okq && okp && tensor_map_encoder() != nullptr) ? 1 : 0;
    }
    a.cfg = *cfg;
    a.prm = *prm;
    if (bat) a.bat = *bat; else memset(&a.bat, 0, sizeof(a.bat));
    a.wrk = *wrk;
    a.tile_batch = asac_sac_tile_batch(cfg);
    if (a.tile_batch < 1) return a.tile_batch;
    a.mode = 0;
    a.late_wait = 0;  // the chains switch it on (chain_late_wait)
    static const int push_combine = [] {
        const char *e = getenv("ASAC_PUSH_COMBINE");
        return e ? atoi(e) : 1;
    }();
    a.push_combine = push_combine;
    a.pi_handoff = 0;
    a.q_sb_handoff = 0;
    const int tiles = (cfg->batch + a.tile_batch - 1) / a.tile_batch;
    ASAC_REQUIRE(wrk->n_tiles == tiles, "work.n_tiles %d != ceil(B / tile_batch) = %d", wrk->n_tiles, tiles);
    ASAC_UNSUPPORTED(tiles > 1024, "batch %d needs %d tiles (> 1024)", cfg->batch, tiles);
    return ASAC_OK;
}

// raises the kernel's dynamic shared-memory limit once per (device, kernel); repeated calls with
// a size already granted do nothing, so CUDA-graph capture never sees an attribute call
template <typename K>
static int set_smem(K kernel, int bytes, const char *name) {
    ASAC_UNSUPPORTED(bytes > kSmemLimit, "%s needs %d bytes of shared memory (> %d)", name, bytes, kSmemLimit);
    if (bytes <= 48 * 1024) return ASAC_OK;
    constexpr int KSLOTS = 12;
    static thread_local int granted[KSLOTS][16];  // [kernel slot][device]
    static thread_local const void *slots[KSLOTS] = {};
    int dev = 0;
    ASAC_CUDA(cudaGetDevice(&dev));
    int slot = 0;
    while (slot < KSLOTS && slots[slot] != nullptr && slots[slot] != (const void *)kernel) ++slot;
    if (slot == KSLOTS || dev >= 16) {
        ASAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        return ASAC_OK;
    }
    slots[slot] = (const void *)kernel;
    if (granted[slot][dev] < bytes) {
        ASAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        granted[slot][dev] = bytes;
    }
    return ASAC_OK;
}

extern "C" int asac_sac_polyak(const AsacSacConfig *cfg, const AsacSacParams *prm, float force_tau, void *stream) {
    int rc = validate(cfg);
    if (rc != ASAC_OK) return rc;
    const int64_t count = net_stride(q_shape(*cfg)) * cfg->ensemble;
    const int force = force_tau >= 0.f;
    k_polyak<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        prm->q_target, prm->q, count, prm->counters, cfg->update_target_per_step, force ? force_tau : cfg->tau,
        force ? (float)(1.0 - (double)force_tau) : cfg->one_minus_tau, force);
    ASAC_LAUNCHED("k_polyak");
    return ASAC_OK;
}

// The tcgen05 layer engine serves the stock width (UMMA_M = hidden = 64) when every row batch of a tile fits
// one MMA (<= 256 rows) and the first layers' K fits the 64-wide operand planes — and when the tile holds more rows
// than ONE FFMA pass (PASS_ROWS): a 16-row tile is latency-bound either way and the FFMA pass is the shorter chain
// (config 2, measured: 32.8 vs 34.8 us for the post pass), from two passes on the tensor-core layer wins
// (config 3: 36.9 -> 26.7 us, config 4: 61.5 -> 49.1 us).  ASAC_TC=0 keeps the FFMA kernels, ASAC_TC=2 forces
// the tensor-core engine for every tile size it supports.
static int tc_mode() {
    static const int mode = [] {
        const char *e = getenv("ASAC_TC");
        return e ? atoi(e) : 1;
    }();
    return mode;
}
static bool tc_enabled() { return tc_mode() != 0; }
static bool value_pass_on_tc(const AsacSacConfig &c, int tile_batch, int mode) {
    if (!tc_enabled() || c.q_hidden != TCF_M || c.pi_hidden != TCF_M || c.state_size + c.action_size > TCF_M) return false;
    const ValueTcPlan p = value_tc_plan(c, tile_batch, mode);
    if (p.RA <= PASS_ROWS && tc_mode() < 2) return false;
    return p.RA <= TCF_MAX_ROWS && p.total * 4 <= 227 * 1024;
}

extern "C" int asac_sac_value_pass_on_tc(const AsacSacConfig *cfg, int mode) {
    if (!cfg || validate(cfg) != ASAC_OK) return 0;
    const int tile = asac_sac_tile_batch(cfg);
    return tile >= 1 && value_pass_on_tc(*cfg, tile, mode) ? 1 : 0;
}

static int launch_value_pass(SacArgs &a, int mode, void *stream) {
    a.mode = mode;
    if (value_pass_on_tc(a.cfg, a.tile_batch, mode)) {
        const ValueTcPlan tp = value_tc_plan(a.cfg, a.tile_batch, mode);
        static_assert(sizeof(ValueTcPlan) <= sizeof(a.plan), "SacArgs::plan too small");
        memcpy(a.plan, &tp, sizeof(tp));
        const int bytes = tp.total * 4;
        int rc = set_smem(k_value_pass_tc, bytes, "k_value_pass_tc");
        if (rc != ASAC_OK) return rc;
        ASAC_CUDA(launch_ex(k_value_pass_tc, dim3(a.wrk.n_tiles, a.cfg.ensemble), dim3(NT), (size_t)bytes,
                            (cudaStream_t)stream, a.cfg.ensemble, true, a));
        ASAC_LAUNCHED("k_value_pass_tc");
        return ASAC_OK;
    }
    const ValuePlan vp = value_plan(a.cfg, a.tile_batch, mode, a.q_sb_handoff);
    static_assert(sizeof(ValuePlan) <= sizeof(a.plan) && sizeof(GradPlan) <= sizeof(a.plan), "SacArgs::plan too small");
    memcpy(a.plan, &vp, sizeof(vp));
    const int bytes = vp.total * 4;
    int rc = set_smem(k_value_pass, bytes, "k_value_pass");
    if (rc != ASAC_OK) return rc;
    ASAC_CUDA(launch_ex(k_value_pass, dim3(a.wrk.n_tiles, a.cfg.ensemble), dim3(NT), (size_t)bytes, (cudaStream_t)stream,
                        a.cfg.ensemble, true, a));
    ASAC_LAUNCHED("k_value_pass");
    return ASAC_OK;
}

extern "C" int asac_sac_target_y(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                                 const AsacSacWork *wrk, void *stream) {
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(bat && bat->states && bat->eps_y, "asac_sac_target_y: missing batch tensors");
    return launch_value_pass(a, 0, stream);
}

extern "C" int asac_sac_post(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                             const AsacSacWork *wrk, void *stream) {
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(bat && bat->states && bat->eps_td && bat->eps_alpha, "asac_sac_post: missing batch tensors");
    return launch_value_pass(a, 1, stream);
}

static int launch_q_backward(SacArgs &a, void *stream) {
    const GradPlan gp = grad_plan(a.cfg, false);
    memcpy(a.plan, &gp, sizeof(gp));
    const int bytes = gp.total * 4;
    int rc = set_smem(k_q_backward, bytes, "k_q_backward");
    if (rc != ASAC_OK) return rc;
    ASAC_CUDA(launch_ex(k_q_backward, dim3(a.wrk.n_tiles, a.cfg.ensemble), dim3(NT), (size_t)bytes, (cudaStream_t)stream,
                        0, true, a));
    ASAC_LAUNCHED("k_q_backward");
    return ASAC_OK;
}

extern "C" int asac_sac_q_backward(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                                   const AsacSacWork *wrk, void *stream) {
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    return launch_q_backward(a, stream);
}

static int launch_policy_backward(SacArgs &a, void *stream) {
    const GradPlan gp = grad_plan(a.cfg, true, a.pi_handoff);
    memcpy(a.plan, &gp, sizeof(gp));
    const int bytes = gp.total * 4;
    int rc = set_smem(k_policy_backward, bytes, "k_policy_backward");
    if (rc != ASAC_OK) return rc;
    ASAC_CUDA(launch_ex(k_policy_backward, dim3(a.wrk.n_tiles, a.cfg.ensemble), dim3(NT), (size_t)bytes,
                        (cudaStream_t)stream, a.cfg.ensemble, true, a));
    ASAC_LAUNCHED("k_policy_backward");
    return ASAC_OK;
}

extern "C" int asac_sac_policy_backward(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                                        const AsacSacWork *wrk, void *stream) {
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    return launch_policy_backward(a, stream);
}

// see SacArgs::pi_handoff.  ASAC_PI_HANDOFF=0 disables it.
static bool pi_handoff(const SacArgs &a) {
    static const int on = [] {
        const char *e = getenv("ASAC_PI_HANDOFF");
        return e ? atoi(e) : 1;
    }();
    const AsacSacConfig &c = a.cfg;
    const NetShape ps = pi_shape(c);
    const int rows = a.tile_batch * (c.seq_len - c.burn_in);  // policy rows of a tile in the train pass
    return on && c.rep_kind == 0 && !value_pass_on_tc(c, a.tile_batch, 0) && rows <= PASS_ROWS &&
           (int64_t)ps.depth * PASS_ROWS * sac_lda(c) <= net_stride(ps);
}

static int chain_late_wait() {
    static const int on = [] {
        const char *e = getenv("ASAC_LATE_WAIT");
        return e ? atoi(e) : 1;
    }();
    return on;
}

// The fused chains hand Q_i(s_b, a_b) from the policy backward to the post pass when a tile's rows leave room for it
// in the critic pass (2 x tile_batch <= 16) and the post pass runs on the FFMA kernel.  ASAC_Q_HANDOFF=0 disables it.
static bool q_sb_handoff(const SacArgs &a, bool need_post) {
    static const int on = [] {
        const char *e = getenv("ASAC_Q_HANDOFF");
        return e ? atoi(e) : 1;
    }();
    return on && need_post && 2 * a.tile_batch <= PASS_ROWS && !value_pass_on_tc(a.cfg, a.tile_batch, 1);
}

// fills the kernel-side exchange descriptor for gradient kind `which` (0 critics, 1 policy, 2 alpha)
static int make_exchange(PeerExchange &x, const AsacSacConfig *cfg, const AsacPeerTable *peers, int which) {
    memset(&x, 0, sizeof(x));
    if (!peers || peers->world <= 1) return ASAC_OK;
    ASAC_REQUIRE(peers->world <= ASAC_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                 "peer table: world %d / rank %d", peers->world, peers->rank);
    const int64_t nq = net_stride(q_shape(*cfg)) * cfg->ensemble, np = net_stride(pi_shape(*cfg));
    x.world = peers->world; x.rank = peers->rank;
    x.total = nq + np + 4 + cfg->rep_param_stride;
    x.off = which == 0 ? 0 : (which == 1 ? nq : (which == 2 ? nq + np : nq + np + 4));  // 3: representation
    ASAC_REQUIRE(peers->recv_words >= 2 * (int64_t)peers->world * x.total,
                 "peer table: receive buffers hold %lld words, need %lld", (long long)peers->recv_words,
                 (long long)(2 * (int64_t)peers->world * x.total));
    for (int i = 0; i < peers->world; ++i) {
        ASAC_REQUIRE(peers->recv[i] && (((uintptr_t)peers->recv[i]) & 7) == 0, "peer table: bad mapping for rank %d", i);
        x.recv[i] = reinterpret_cast<unsigned long long *>(peers->recv[i]);
    }
    return ASAC_OK;
}

extern "C" int64_t asac_peer_recv_words(const AsacSacConfig *cfg, int world) {
    if (validate(cfg) != ASAC_OK) return -1;
    return 2 * (int64_t)world *
           (net_stride(q_shape(*cfg)) * cfg->ensemble + net_stride(pi_shape(*cfg)) + 4 + cfg->rep_param_stride);
}

extern "C" int asac_peer_timeouts(int reset) {
    unsigned int n = 0;
    if (cudaMemcpyFromSymbol(&n, asac::g_peer_timeouts, sizeof(n)) != cudaSuccess) return -1;
    if (reset && n) {
        const unsigned int zero = 0;
        cudaMemcpyToSymbol(asac::g_peer_timeouts, &zero, sizeof(zero));
    }
    return (int)n;
}

// (Measured and dropped, round 2: the same step on <= 20 fat CTAs launched as 2-CTA clusters, so that the 64 clusters
// of the kernel behind it find 64 free TPCs and run their parameter-independent part during the optimiser step.  The
// overlap worked — every policy-backward CTA started 1.4 us after the critic backward's exit — but 18 CTAs take 10.4 us
// for what this grid does in 6.2, longer than the part it hides behind: 151.0 vs 149.6 us per step.)
static int launch_adam_kernel(const AdamArgs &a, void *stream) {
    ASAC_CUDA(launch_ex(k_reduce_adam, dim3((unsigned)((a.count + ADAM_PARAMS_PER_CTA - 1) / ADAM_PARAMS_PER_CTA)),
                        dim3(ADAM_PARAMS_PER_CTA * ADAM_TILE_GROUPS), 0, (cudaStream_t)stream, 0, true, a));
    ASAC_LAUNCHED("k_reduce_adam");
    return ASAC_OK;
}

static int launch_reduce_adam(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk, int which,
                              int do_reduce, int do_adam, float grad_scale, void *stream, int do_td = 0,
                              const AsacPeerTable *peers = nullptr) {
    int rc = validate(cfg);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(which >= 0 && which <= 2, "which must be 0 (critics), 1 (policy) or 2 (alpha)");
    if (which == 2) {
        ASAC_UNSUPPORTED(wrk->n_tiles > 1024, "n_tiles %d > 1024", wrk->n_tiles);
        const int threads = cfg->batch >= 1024 ? 1024 : ((cfg->batch + 31) / 32) * 32;
        k_alpha_td<<<1, threads, 0, (cudaStream_t)stream>>>(*prm, *wrk, wrk->n_tiles, cfg->batch, cfg->ensemble,
                                                            do_reduce, do_adam, do_td, grad_scale,
                                                            cfg->learning_rate);
        ASAC_LAUNCHED("k_alpha_td");
        return ASAC_OK;
    }
    AdamArgs a;
    if ((rc = make_exchange(a.px, cfg, peers, which)) != ASAC_OK) return rc;
    if (which == 0) {
        const int64_t stride = net_stride(q_shape(*cfg));
        a.param = prm->q; a.m = prm->q_m; a.v = prm->q_v;
        a.part = do_reduce ? wrk->grad_q_part : nullptr;
        a.grad = wrk->grad_q;
        a.step = prm->counters + 1;
        a.count = stride * cfg->ensemble;
        a.tile_stride = a.count;
    } else {
        const int64_t stride = net_stride(pi_shape(*cfg));
        a.param = prm->pi; a.m = prm->pi_m; a.v = prm->pi_v;
        a.part = do_reduce ? wrk->grad_pi_part : nullptr;
        a.grad = wrk->grad_pi;
        a.step = prm->counters + 2;
        a.count = stride;
        a.tile_stride = stride;
    }
    a.n_tiles = wrk->n_tiles;
    a.write_grad = 1;
    a.do_adam = do_adam;
    a.grad_scale = grad_scale;
    a.lr = cfg->learning_rate;
    return launch_adam_kernel(a, stream);
}

static int bump(const AsacSacParams *prm, int mask, void *stream) {
    k_bump<<<1, 32, 0, (cudaStream_t)stream>>>(prm->counters, mask);
    ASAC_LAUNCHED("k_bump");
    return ASAC_OK;
}

extern "C" int asac_sac_reduce_grads(const AsacSacConfig *cfg, const AsacSacWork *wrk, int which, void *stream) {
    AsacSacParams none;
    memset(&none, 0, sizeof(none));
    return launch_reduce_adam(cfg, &none, wrk, which, 1, 0, 1.f, stream);
}

extern "C" int asac_sac_adam(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk, int which,
                             float grad_scale, void *stream) {
    int rc = launch_reduce_adam(cfg, prm, wrk, which, 0, 1, grad_scale, stream);
    if (rc != ASAC_OK) return rc;
    return bump(prm, 1 << (which + 1), stream);
}

extern "C" int asac_sac_reduce_adam(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk,
                                    int which, void *stream) {
    int rc = launch_reduce_adam(cfg, prm, wrk, which, 1, 1, 1.f, stream, which == 2 ? 1 : 0);
    if (rc != ASAC_OK) return rc;
    return bump(prm, 1 << (which + 1), stream);
}

extern "C" int asac_sac_advance_step(const AsacSacParams *prm, void *stream) { return bump(prm, 1, stream); }

extern "C" int asac_sac_td_error(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk,
                                 void *stream) {
    return launch_reduce_adam(cfg, prm, wrk, 2, 0, 0, 1.f, stream, 1);
}

extern "C" int asac_sac_step_networks(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                                      const AsacSacWork *wrk, int with_polyak, const AsacPeerTable *peers,
                                      void *stream) {
    const float gscale = (peers && peers->world > 1) ? 1.f / (float)peers->world : 1.f;
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    a.late_wait = chain_late_wait();  // this function launches the whole chain: the kernels know their predecessors
    a.pi_handoff = pi_handoff(a) ? 1 : 0;
    ASAC_REQUIRE(bat && bat->states && bat->eps_y && bat->eps_pi, "asac_sac_step: missing batch tensors");
    if (with_polyak && (rc = asac_sac_polyak(cfg, prm, -1.f, stream)) != ASAC_OK) return rc;
    if ((rc = launch_value_pass(a, 0, stream)) != ASAC_OK) return rc;
    if ((rc = launch_q_backward(a, stream)) != ASAC_OK) return rc;
    if ((rc = launch_reduce_adam(cfg, prm, wrk, 0, 1, 1, gscale, stream, 0, peers)) != ASAC_OK) return rc;
    const bool need_post = cfg->use_auto_alpha || cfg->use_n_step_is || cfg->use_priority;
    a.q_sb_handoff = q_sb_handoff(a, need_post) ? 1 : 0;
    if ((rc = launch_policy_backward(a, stream)) != ASAC_OK) return rc;
    if ((rc = launch_reduce_adam(cfg, prm, wrk, 1, 1, 1, gscale, stream, 0, peers)) != ASAC_OK) return rc;
    if (need_post) {
        ASAC_REQUIRE(bat->eps_td && bat->eps_alpha, "asac_sac_step: missing eps_td / eps_alpha");
        if ((rc = launch_value_pass(a, 1, stream)) != ASAC_OK) return rc;
    }
    return ASAC_OK;
}

static int flat_reduce_adam(float *param, float *m, float *v, const float *grad_part, int n_tiles, int64_t tile_stride,
                            int64_t count, float *grad, const int64_t *step_counter, double learning_rate,
                            const PeerExchange *px, float grad_scale, void *stream) {
    ASAC_REQUIRE(param && m && v && grad_part && grad && step_counter && n_tiles > 0 && count > 0,
                 "asac_flat_reduce_adam: bad arguments");
    AdamArgs a;
    if (px) a.px = *px; else memset(&a.px, 0, sizeof(a.px));
    a.param = param; a.m = m; a.v = v; a.part = grad_part; a.grad = grad; a.step = step_counter;
    a.count = count; a.tile_stride = tile_stride; a.n_tiles = n_tiles;
    a.write_grad = 1; a.do_adam = 1; a.grad_scale = grad_scale; a.lr = learning_rate;
    return launch_adam_kernel(a, stream);
}

extern "C" int asac_flat_reduce_adam(float *param, float *m, float *v, const float *grad_part, int n_tiles,
                                     int64_t tile_stride, int64_t count, float *grad, const int64_t *step_counter,
                                     double learning_rate, void *stream) {
    return flat_reduce_adam(param, m, v, grad_part, n_tiles, tile_stride, count, grad, step_counter, learning_rate,
                            nullptr, 1.f, stream);
}

extern "C" int asac_flat_polyak(float *target, const float *source, int64_t count, const int64_t *counters,
                                int per_step, float tau, float one_minus_tau, int force, void *stream) {
    ASAC_REQUIRE(target && source && count > 0 && (force || (counters && per_step >= 1)), "asac_flat_polyak: bad arguments");
    k_polyak<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(target, source, count, counters,
                                                                               per_step, tau, one_minus_tau, force);
    ASAC_LAUNCHED("k_polyak");
    return ASAC_OK;
}

extern "C" int asac_sac_step_networks_rep(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                                          const AsacSacWork *wrk, const AsacGruRep *rep, int with_polyak,
                                          const AsacPeerTable *peers, void *stream) {
    const float gscale = (peers && peers->world > 1) ? 1.f / (float)peers->world : 1.f;
    SacArgs a;
    int rc = make_args(a, cfg, prm, bat, wrk);
    if (rc != ASAC_OK) return rc;
    a.late_wait = chain_late_wait();  // this function launches the whole chain: the kernels know their predecessors
    ASAC_REQUIRE(rep && cfg->rep_kind == 1, "asac_sac_step_networks_rep: cfg.rep_kind must be 1 (GRU)");
    ASAC_REQUIRE(rep->shape.hidden == cfg->state_size && rep->shape.action_size == cfg->action_size,
                 "asac_sac_step_networks_rep: GRU width %d / action %d vs state_size %d / action_size %d",
                 rep->shape.hidden, rep->shape.action_size, cfg->state_size, cfg->action_size);
    ASAC_REQUIRE(bat && bat->eps_y && bat->eps_pi && bat->actions, "asac_sac_step: missing batch tensors");
    ASAC_REQUIRE(bat->states == rep->states && bat->states_post == rep->states_post &&
                     bat->target_states == rep->target_states && wrk->grad_state,
                 "asac_sac_step_networks_rep: batch / work do not point at the representation's buffers");
    ASAC_REQUIRE(rep->params && rep->params_target && rep->m && rep->v && rep->obs && rep->hn && rep->hn_post &&
                     rep->save && rep->grad_part && rep->grad, "asac_sac_step_networks_rep: null pointer in rep");
    const int tile = asac_gru_backward_tile(&rep->shape, cfg->burn_in);
    if (tile < 1) return tile;
    ASAC_REQUIRE(rep->rep_tiles == cfg->batch, "rep.rep_tiles %d != batch %d", rep->rep_tiles, cfg->batch);
    const int64_t P = asac_gru_param_count(&rep->shape), Ps = (P + 3) / 4 * 4;
    const int B = cfg->batch, L = cfg->seq_len;
    PeerExchange px_rep;
    ASAC_REQUIRE(!(peers && peers->world > 1) || cfg->rep_param_stride == Ps,
                 "cfg.rep_param_stride %d != the representation's gradient stride %lld", cfg->rep_param_stride, (long long)Ps);
    if ((rc = make_exchange(px_rep, cfg, peers, 3)) != ASAC_OK) return rc;
    if (with_polyak) {
        if ((rc = asac_sac_polyak(cfg, prm, -1.f, stream)) != ASAC_OK) return rc;
        if ((rc = asac_flat_polyak(rep->params_target, rep->params, P, prm->counters, cfg->update_target_per_step,
                                   cfg->tau, cfg->one_minus_tau, 0, stream)) != ASAC_OK) return rc;
    }
    // get_l_states x 2 (sac_base.py:2066-2078): online (gates kept for the backward pass) and target
    AsacGruNet nets[2] = {{rep->params, rep->states, rep->hn, rep->save},
                          {rep->params_target, rep->target_states, nullptr, nullptr}};
    if ((rc = asac_gru_forward(&rep->shape, nets, 2, rep->obs, bat->actions, cfg->bn_stride, nullptr, rep->h0,
                               rep->h0_b_stride, B, L, stream)) != ASAC_OK) return rc;
    if ((rc = launch_value_pass(a, 0, stream)) != ASAC_OK) return rc;
    if ((rc = launch_q_backward(a, stream)) != ASAC_OK) return rc;
    if ((rc = launch_reduce_adam(cfg, prm, wrk, 0, 1, 1, gscale, stream, 0, peers)) != ASAC_OK) return rc;
    // the representation's share of loss.backward() and optimizer_rep.step() (sac_base.py:1573-1601)
    if ((rc = asac_gru_backward(&rep->shape, rep->params, rep->obs, bat->actions, cfg->bn_stride, nullptr, rep->h0,
                                rep->h0_b_stride, B, L, cfg->burn_in, wrk->grad_state, cfg->ensemble, rep->hn,
                                rep->save, rep->grad_part, stream)) != ASAC_OK) return rc;
    if ((rc = flat_reduce_adam(rep->params, rep->m, rep->v, rep->grad_part, rep->rep_tiles, Ps, P, rep->grad,
                               prm->counters + 4, cfg->learning_rate, &px_rep, gscale, stream)) != ASAC_OK) return rc;
    // get_l_states again with the new weights (sac_base.py:2099-2105)
    AsacGruNet again = {rep->params, rep->states_post, rep->hn_post, nullptr};
    if ((rc = asac_gru_forward(&rep->shape, &again, 1, rep->obs, bat->actions, cfg->bn_stride, nullptr, rep->h0,
                               rep->h0_b_stride, B, L, stream)) != ASAC_OK) return rc;
    const bool need_post = cfg->use_auto_alpha || cfg->use_n_step_is || cfg->use_priority;
    a.q_sb_handoff = q_sb_handoff(a, need_post) ? 1 : 0;
    if ((rc = launch_policy_backward(a, stream)) != ASAC_OK) return rc;
    if ((rc = launch_reduce_adam(cfg, prm, wrk, 1, 1, 1, gscale, stream, 0, peers)) != ASAC_OK) return rc;
    if (need_post) {
        ASAC_REQUIRE(bat->eps_td && bat->eps_alpha, "asac_sac_step: missing eps_td / eps_alpha");
        if ((rc = launch_value_pass(a, 1, stream)) != ASAC_OK) return rc;
    }
    return ASAC_OK;
}

extern "C" int asac_sac_staged_tail(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk,
                                    void *stream) {
    int rc;
    int mask = 1 | 2 | 4 | (cfg->rep_kind != 0 ? 16 : 0);
    const bool need_post = cfg->use_auto_alpha || cfg->use_n_step_is || cfg->use_priority;
    if (cfg->use_auto_alpha || need_post) {
        const int aa = cfg->use_auto_alpha ? 1 : 0;
        if ((rc = launch_reduce_adam(cfg, prm, wrk, 2, aa, aa, 1.f, stream, need_post ? 1 : 0)) != ASAC_OK) return rc;
        if (aa) mask |= 8;
    }
    return bump(prm, mask, stream);
}

extern "C" int asac_sac_step(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacBatch *bat,
                             const AsacSacWork *wrk, void *stream) {
    int rc = asac_sac_step_networks(cfg, prm, bat, wrk, 1, nullptr, stream);
    if (rc != ASAC_OK) return rc;
    int mask = 1 | 2 | 4;
    const bool need_post = cfg->use_auto_alpha || cfg->use_n_step_is || cfg->use_priority;
    if (cfg->use_auto_alpha || need_post) {
        const int aa = cfg->use_auto_alpha ? 1 : 0;
        if ((rc = launch_reduce_adam(cfg, prm, wrk, 2, aa, aa, 1.f, stream, need_post ? 1 : 0)) != ASAC_OK) return rc;
        if (aa) mask |= 8;
    }
    return bump(prm, mask, stream);
}

extern "C" int asac_sac_finish_step(const AsacSacConfig *cfg, const AsacSacParams *prm, const AsacSacWork *wrk,
                                    float *nodes, int64_t capacity, const int64_t *store_ids,
                                    const int64_t *data_ids, double *per_state, const AsacPeerTable *peers,
                                    void *stream) {
    int rc = validate(cfg);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(is_pow2(capacity), "asac_sac_finish_step: capacity is not a power of two");
    ASAC_REQUIRE(cfg->batch <= 1024, "asac_sac_finish_step: batch %d > 1024", cfg->batch);
    ASAC_REQUIRE(prm && wrk && store_ids && data_ids && per_state, "asac_sac_finish_step: null pointer");
    EpilogueArgs a;
    if ((rc = make_exchange(a.px, cfg, peers, 2)) != ASAC_OK) return rc;
    a.grad_scale = (peers && peers->world > 1) ? 1.f / (float)peers->world : 1.f;
    a.prm = *prm; a.wrk = *wrk;
    a.n_tiles = wrk->n_tiles; a.batch = cfg->batch; a.ensemble = cfg->ensemble;
    a.use_auto_alpha = cfg->use_auto_alpha ? 1 : 0;
    a.counter_mask = 1 | 2 | 4 | (cfg->use_auto_alpha ? 8 : 0) | (cfg->rep_kind != 0 ? 16 : 0);
    a.lr = cfg->learning_rate;
    a.nodes = nodes; a.capacity = capacity; a.levels = tree_levels(capacity);
    a.store_ids = store_ids; a.data_ids = data_ids;
    a.td_min = cfg->td_error_min; a.td_max = cfg->td_error_max; a.per_alpha = cfg->per_alpha;
    a.per_state = per_state;
    const int threads = ((cfg->batch + 31) / 32) * 32;
    ASAC_CUDA(launch_ex(k_step_epilogue, dim3(1), dim3(threads), 0, (cudaStream_t)stream, 0, true, a));
    ASAC_LAUNCHED("k_step_epilogue");
    return ASAC_OK;
}

extern "C" int asac_l2_prefetch(const void *const *regions, const int64_t *bytes, int n, void *stream) {
    ASAC_REQUIRE(regions && bytes && n >= 1 && n <= 8, "asac_l2_prefetch: 1..8 regions");
    PrefetchArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    for (int i = 0; i < n; ++i) {
        ASAC_REQUIRE(regions[i] && bytes[i] > 0, "asac_l2_prefetch: empty region %d", i);
        a.base[i] = reinterpret_cast<const char *>(regions[i]);
        a.lines[i] = (bytes[i] + 127) / 128;
        a.first[i + 1] = a.first[i] + a.lines[i];
    }
    k_l2_prefetch<<<(unsigned)((a.first[n] + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_l2_prefetch");
    return ASAC_OK;
}

extern "C" int asac_fill_normal(float *out, int64_t n, uint64_t seed, const int64_t *counter, int stream_id,
                                void *stream) {
    if (n <= 0) return ASAC_OK;
    const int64_t quads = (n + 3) / 4;
    k_fill_normal<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, counter, stream_id);
    ASAC_LAUNCHED("k_fill_normal");
    return ASAC_OK;
}

extern "C" int asac_mlp_forward(const float *params, int in_dim, int hidden, int depth, int out_dim, const float *x,
                                int64_t rows, float *out, void *stream) {
    ASAC_UNSUPPORTED(!(hidden == 16 || hidden == 32 || hidden == 64 || hidden == 128), "hidden width %d", hidden);
    ASAC_UNSUPPORTED(depth < 1 || depth > ASAC_MAX_DEPTH, "depth %d", depth);
    ASAC_REQUIRE(in_dim > 0 && out_dim > 0 && rows > 0, "asac_mlp_forward: bad sizes");
    MlpArgs a;
    a.params = params; a.x = x; a.out = out;
    a.s = NetShape{in_dim, hidden, depth, out_dim};
    a.rows = rows;
    a.rows_per_cta = 32;
    const int bytes = mlp_plan(a.s, a.rows_per_cta).total * 4;
    int rc = set_smem(k_mlp_forward, bytes, "k_mlp_forward");
    if (rc != ASAC_OK) return rc;
    k_mlp_forward<<<(unsigned)((rows + a.rows_per_cta - 1) / a.rows_per_cta), NT, bytes, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_mlp_forward");
    return ASAC_OK;
}

extern "C" int asac_mlp_forward_tc(const float *params, int in_dim, int hidden, int depth, int out_dim, const float *x,
                                   int64_t rows, float *out, void *stream);

extern "C" int asac_policy_act(const float *params, int state_size, int hidden, int depth, int action_size,
                               const float *states, int64_t rows, const float *eps, const float *offline_action,
                               int disable_sample, uint64_t seed, const int64_t *counter, float *scratch,
                               float *out_action, float *out_prob, int use_tensor_cores, void *stream) {
    ASAC_REQUIRE(params && states && scratch && out_action && out_prob, "asac_policy_act: null pointer");
    ASAC_REQUIRE(rows > 0 && action_size > 0 && action_size <= 64, "asac_policy_act: bad sizes");
    ASAC_REQUIRE(eps || offline_action || disable_sample || counter, "asac_policy_act: need eps or a Philox counter");
    int rc;
    if (use_tensor_cores && hidden == 64 && 2 * action_size <= 16)
        rc = asac_mlp_forward_tc(params, state_size, hidden, depth, 2 * action_size, states, rows, scratch, stream);
    else
        rc = asac_mlp_forward(params, state_size, hidden, depth, 2 * action_size, states, rows, scratch, stream);
    if (rc != ASAC_OK) return rc;
    ActArgs a;
    a.pre = scratch; a.eps = eps; a.offline = offline_action; a.action = out_action; a.prob = out_prob;
    a.rows = rows; a.A = action_size; a.disable_sample = disable_sample; a.seed = seed; a.counter = counter;
    k_policy_sample<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_policy_sample");
    return ASAC_OK;
}

// debug: phase clocks of CTA (0,0) of the last value pass / critic backward / policy backward launch
// debug (-DASAC_PROBES): reads and resets the %globaltimer stamps (max slots to 0, min slots to ~0)
extern "C" int asac_debug_global_stamps(uint64_t *out_host) {
    ASAC_CUDA(cudaMemcpyFromSymbol(out_host, g_gt, sizeof(unsigned long long) * 8));
    unsigned long long init[8] = {0, ~0ull, 0, 0, ~0ull, 0, 0, 0};
    ASAC_CUDA(cudaMemcpyToSymbol(g_gt, init, sizeof(init)));
    return ASAC_OK;
}

extern "C" int asac_debug_phase_clocks(int64_t *out_host) {
    ASAC_REQUIRE(out_host != nullptr, "asac_debug_phase_clocks: null pointer");
    ASAC_CUDA(cudaMemcpyFromSymbol(out_host, g_phase_clock, sizeof(long long) * 3 * 32));
    // slots [0][29], [0][30]: cycles CTA (0,0) waited for staged weights since the last call, number of waits
    long long w[2] = {0, 0}, zero[2] = {0, 0};
    ASAC_CUDA(cudaMemcpyFromSymbol(w, g_pipe_wait, sizeof(w)));
    ASAC_CUDA(cudaMemcpyToSymbol(g_pipe_wait, zero, sizeof(zero)));
    out_host[29] = w[0];
    out_host[30] = w[1];
    // slots [1][20..24]: layer_forward segments (GEMM, barrier, epilogue, barrier, passes)
    long long seg[8], zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ASAC_CUDA(cudaMemcpyFromSymbol(seg, g_layer_seg, sizeof(seg)));
    ASAC_CUDA(cudaMemcpyToSymbol(g_layer_seg, zero8, sizeof(zero8)));
    for (int i = 0; i < 5; ++i) out_host[32 + 20 + i] = seg[i];
    return ASAC_OK;
}

// ------------------------------------------------------------------------------------ discrete action branches
static int fill_dnets(DNets &n, const AsacDiscreteConfig *d, const float *params, int64_t member_stride, int members) {
    ASAC_REQUIRE(d && params && members >= 1 && members <= ASAC_MAX_ENSEMBLE, "discrete nets: bad arguments");
    ASAC_UNSUPPORTED(d->branches < 1 || d->branches > ASAC_MAX_BRANCHES, "%d discrete action branches (1..%d)", d->branches,
                     ASAC_MAX_BRANCHES);
    ASAC_UNSUPPORTED(!(d->hidden == 16 || d->hidden == 32 || d->hidden == 64 || d->hidden == 128),
                     "d_dense_n %d not in {16, 32, 64, 128}", d->hidden);
    ASAC_UNSUPPORTED(d->depth < 1 || d->depth > ASAC_MAX_DEPTH, "d_dense_depth %d outside [1, %d]", d->depth, ASAC_MAX_DEPTH);
    ASAC_REQUIRE(d->state_size > 0, "discrete nets: state_size");
    memset(&n, 0, sizeof(n));
    n.params = params; n.member_stride = member_stride; n.members = members; n.branches = d->branches;
    n.S = d->state_size; n.H = d->hidden; n.depth = d->depth;
    int64_t off = 0;
    int col = 0;
    for (int k = 0; k < d->branches; ++k) {
        ASAC_REQUIRE(d->sizes[k] >= 1, "discrete branch %d has size %d", k, d->sizes[k]);
        n.sizes[k] = d->sizes[k];
        n.branch_off[k] = off;
        n.col_off[k] = col;
        off += net_stride(NetShape{n.S, n.H, n.depth, d->sizes[k]});
        col += d->sizes[k];
    }
    n.D = col;
    ASAC_UNSUPPORTED(col > D_MAX_COLS, "%d discrete action columns > %d", col, D_MAX_COLS);
    return ASAC_OK;
}

extern "C" int64_t asac_dnets_member_floats(const AsacDiscreteConfig *d) {
    if (!d) return 0;
    int64_t off = 0;
    for (int k = 0; k < d->branches && k < ASAC_MAX_BRANCHES; ++k)
        off += net_stride(NetShape{d->state_size, d->hidden, d->depth, d->sizes[k]});
    return off;
}

extern "C" int asac_dnets_tiles(int rows) { return (rows + PASS_ROWS - 1) / PASS_ROWS; }

extern "C" int asac_dnets_forward(const AsacDiscreteConfig *d, const float *params, int64_t member_stride, int members,
                                  const float *x, int64_t x_row_stride, int rows, float *out, void *stream) {
    DFwdArgs a;
    int rc = fill_dnets(a.nets, d, params, member_stride, members);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(x && out && rows > 0 && x_row_stride >= d->state_size, "asac_dnets_forward: bad arguments");
    a.x = x; a.x_row_stride = x_row_stride; a.rows = rows; a.out = out;
    const int bytes = dfwd_plan(a.nets.S, a.nets.H, a.nets.depth, D_MAX_COLS).total * 4;
    rc = set_smem(k_dnets_forward, bytes, "k_dnets_forward");
    if (rc != ASAC_OK) return rc;
    k_dnets_forward<<<dim3(asac_dnets_tiles(rows), members * d->branches), NT, bytes, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_dnets_forward");
    return ASAC_OK;
}

extern "C" int asac_dnets_backward(const AsacDiscreteConfig *d, const float *params, int64_t member_stride, int members,
                                   const float *x, int64_t x_row_stride, int rows, const float *d_out, float *grad_part,
                                   float *d_x, void *stream) {
    DBwdArgs a;
    int rc = fill_dnets(a.nets, d, params, member_stride, members);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(x && d_out && grad_part && rows > 0 && x_row_stride >= d->state_size, "asac_dnets_backward: bad arguments");
    a.x = x; a.x_row_stride = x_row_stride; a.rows = rows; a.d_out = d_out; a.grad_part = grad_part; a.d_x = d_x;
    a.member_floats = asac_dnets_member_floats(d);
    const int bytes = dbwd_plan(a.nets.S, a.nets.H, a.nets.depth, D_MAX_COLS).total * 4;
    rc = set_smem(k_dnets_backward, bytes, "k_dnets_backward");
    if (rc != ASAC_OK) return rc;
    k_dnets_backward<<<dim3(asac_dnets_tiles(rows), members * d->branches), NT, bytes, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_dnets_backward");
    return ASAC_OK;
}

static int d_sizes(const AsacDiscreteConfig *d, int *sizes, int &D) {
    ASAC_REQUIRE(d && d->branches >= 1 && d->branches <= ASAC_MAX_BRANCHES, "discrete config: branches");
    D = 0;
    for (int k = 0; k < ASAC_MAX_BRANCHES; ++k) {
        sizes[k] = k < d->branches ? d->sizes[k] : 0;
        D += sizes[k];
    }
    ASAC_UNSUPPORTED(D > D_MAX_COLS, "%d discrete action columns > %d", D, D_MAX_COLS);
    return ASAC_OK;
}

extern "C" int asac_d_target(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *pi_logits, const float *tq,
                             const float *actions_full, const float *mu_full, const float *pi_probs_d,
                             const float *rewards, const uint8_t *dones, const uint8_t *last_masks,
                             const uint8_t *padding_masks, const float *log_d_alpha, float *d_y, void *stream) {
    DTargetArgs a;
    memset(&a, 0, sizeof(a));
    int rc = d_sizes(d, a.sizes, a.D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && pi_logits && tq && actions_full && rewards && dones && last_masks && padding_masks && log_d_alpha && d_y,
                 "asac_d_target: null pointer");
    ASAC_REQUIRE(!cfg->use_n_step_is || mu_full || pi_probs_d, "asac_d_target: use_n_step_is needs mu probabilities");
    ASAC_UNSUPPORTED(cfg->n_step > ASAC_MAX_NSTEP, "n_step %d > %d", cfg->n_step, ASAC_MAX_NSTEP);
    a.cfg = *cfg; a.branches = d->branches; a.AF = a.D + cfg->action_size;
    a.pi_logits = pi_logits; a.tq = tq; a.actions_full = actions_full; a.mu_full = mu_full; a.pi_probs_d = pi_probs_d;
    a.rewards = rewards; a.dones = dones; a.last_masks = last_masks; a.padding_masks = padding_masks;
    a.log_d_alpha = log_d_alpha; a.post = pi_probs_d ? 1 : 0; a.d_y = d_y;
    k_d_target<<<(cfg->batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_target");
    return ASAC_OK;
}

extern "C" int asac_d_target_dqn(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *eval_q, const float *tq,
                                 const int32_t *perm_target, const int32_t *perm_online, const float *rewards,
                                 const uint8_t *dones, const uint8_t *last_masks, const uint8_t *padding_masks,
                                 float *d_y, void *stream) {
    DDqnArgs a;
    memset(&a, 0, sizeof(a));
    int rc = d_sizes(d, a.sizes, a.D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && eval_q && tq && rewards && dones && last_masks && padding_masks && d_y, "asac_d_target_dqn: null pointer");
    a.cfg = *cfg; a.branches = d->branches; a.eval_q = eval_q; a.tq = tq;
    a.perm_target = perm_target; a.perm_online = perm_online;
    a.Es = (cfg->ensemble_sample > 0 && cfg->ensemble_sample < cfg->ensemble) ? cfg->ensemble_sample : cfg->ensemble;
    a.rewards = rewards; a.dones = dones; a.last_masks = last_masks; a.padding_masks = padding_masks; a.d_y = d_y;
    k_d_target_dqn<<<(cfg->batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_target_dqn");
    return ASAC_OK;
}

extern "C" int asac_d_q_grad(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *q, const float *actions_full,
                             const float *d_y, const float *weights, float scale, float *d_out, float *loss, float *q_single,
                             void *stream) {
    int sizes[ASAC_MAX_BRANCHES], D;
    int rc = d_sizes(d, sizes, D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && q && actions_full && d_y && d_out && loss && q_single, "asac_d_q_grad: null pointer");
    DQGradArgs a{cfg->batch, cfg->seq_len, cfg->burn_in, cfg->ensemble, d->branches, D, D + cfg->action_size,
                 q, actions_full, d_y, weights, scale, d_out, loss, q_single};
    k_d_q_grad<<<(cfg->ensemble * cfg->batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_q_grad");
    return ASAC_OK;
}

extern "C" int asac_d_pi_grad(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, int stride_rows,
                              int row_off, const float *q, const float *mu_full, const float *log_d_alpha, float *d_out,
                              float *loss, float *entropy, void *stream) {
    DPiGradArgs a;
    memset(&a, 0, sizeof(a));
    int rc = d_sizes(d, a.sizes, a.D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && logits && q && mu_full && log_d_alpha && d_out && loss && entropy, "asac_d_pi_grad: null pointer");
    a.B = cfg->batch; a.L = cfg->seq_len; a.b = cfg->burn_in; a.E = cfg->ensemble; a.branches = d->branches;
    a.AF = a.D + cfg->action_size; a.stride_rows = stride_rows; a.row_off = row_off;
    a.logits = logits; a.q = q; a.mu_full = mu_full; a.log_d_alpha = log_d_alpha; a.penalty = d->entropy_penalty;
    a.d_out = d_out; a.loss = loss; a.entropy = entropy;
    k_d_pi_grad<<<(cfg->batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_pi_grad");
    return ASAC_OK;
}

extern "C" int asac_d_probs(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, float *pi_probs_d,
                            float *pi_probs_full, void *stream) {
    DPostArgs a;
    memset(&a, 0, sizeof(a));
    int rc = d_sizes(d, a.sizes, a.D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && logits && pi_probs_d, "asac_d_probs: null pointer");
    a.B = cfg->batch; a.L = cfg->seq_len; a.b = cfg->burn_in; a.E = cfg->ensemble; a.branches = d->branches;
    a.AF = a.D + cfg->action_size; a.logits = logits; a.pi_probs_d = pi_probs_d; a.pi_probs_full = pi_probs_full;
    k_d_probs<<<(cfg->batch * (cfg->seq_len - 1) + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_probs");
    return ASAC_OK;
}

extern "C" int asac_d_alpha(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *logits, float *log_d_alpha,
                            float *m, float *v, const int64_t *step, float grad_scale, float *grad_out, float *loss_out,
                            void *stream) {
    DAlphaArgs a;
    memset(&a, 0, sizeof(a));
    int rc = d_sizes(d, a.sizes, a.D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && logits && log_d_alpha && m && v && step && grad_out && loss_out, "asac_d_alpha: null pointer");
    a.B = cfg->batch; a.L = cfg->seq_len; a.b = cfg->burn_in; a.branches = d->branches;
    a.target_ratio = d->target_d_alpha; a.logits = logits; a.log_d_alpha = log_d_alpha; a.m = m; a.v = v; a.step = step;
    a.lr = cfg->learning_rate; a.grad_scale = grad_scale; a.grad_out = grad_out; a.loss_out = loss_out;
    k_d_alpha<<<1, NT, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_alpha");
    return ASAC_OK;
}

extern "C" int asac_d_td(const AsacSacConfig *cfg, const AsacDiscreteConfig *d, const float *q, const float *actions_full,
                         const float *d_y, float *td_error, int accumulate, void *stream) {
    int sizes[ASAC_MAX_BRANCHES], D;
    int rc = d_sizes(d, sizes, D);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(cfg && q && actions_full && d_y && td_error, "asac_d_td: null pointer");
    DTdArgs a{cfg->batch, cfg->seq_len, cfg->burn_in, cfg->ensemble, d->branches, D, D + cfg->action_size, accumulate,
              q, actions_full, d_y, td_error};
    k_d_td<<<(cfg->batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_d_td");
    return ASAC_OK;
}

extern "C" int asac_ensemble_perms(int32_t *out, int n_perms, int ensemble, uint64_t seed, const int64_t *counter,
                                   void *stream) {
    ASAC_REQUIRE(out && counter && n_perms >= 1 && n_perms <= 64 && ensemble >= 1 && ensemble <= ASAC_MAX_ENSEMBLE,
                 "asac_ensemble_perms: bad arguments");
    k_ensemble_perms<<<1, 64, 0, (cudaStream_t)stream>>>(out, n_perms, ensemble, seed, counter);
    ASAC_LAUNCHED("k_ensemble_perms");
    return ASAC_OK;
}

extern "C" int asac_bump_counters(int64_t *counters, int mask, void *stream) {
    ASAC_REQUIRE(counters, "asac_bump_counters: null pointer");
    k_bump<<<1, 32, 0, (cudaStream_t)stream>>>(counters, mask);
    ASAC_LAUNCHED("k_bump");
    return ASAC_OK;
}
