// Discrete (and the discrete half of hybrid) action branches of the SAC update.
// Included by sac.cu inside namespace asac (uses its plans and the row-tile routines of mlp_tile.cuh).
//
// Reference: ModelQ / ModelPolicy keep one LinearLayers(state -> d_dense_n x d_dense_depth -> d_action_size_k) per
// action branch (nn_models/q.py:60-64, 74-78; policy.py:143-147, 152-160); the losses are
//   _get_y discrete branch        sac_base.py:1356-1421   (mean over the ensemble, expectation over the policy)
//   _train_rep_q                   sac_base.py:1516-1570   (q_single = sum(onehot * q) / branches)
//   _train_policy + entropy penalty sac_base.py:1858-1880
//   _train_alpha                   sac_base.py:1924-1929
//   get_l_probs / _get_td_error    sac_base.py:1159-1189, 2182-2245
//   JointOneHotCategorical         nn_models/policy.py:47-84
// The discrete nets share nothing with the continuous ones but the per-sample sum of the losses, so they run as
// their own launches next to the fused continuous kernels:
//   k_dnets_forward   — K branch nets x M members over 16-row tiles (forward only)
//   k_dnets_backward  — the same walk with saved activations, then the backward pass from a given d loss / d output
//   k_d_*             — the per-row loss math between them (softmax, expectation, V-trace, gradients of the heads)
#pragma once

constexpr int D_MAX_COLS = 64;  // sum of the branch sizes handled by the per-row kernels

struct DNets {
    const float *params;      // member 0, branch 0
    int64_t member_stride;    // floats between members (0 for a single net set)
    int members, branches;
    int sizes[ASAC_MAX_BRANCHES];
    int64_t branch_off[ASAC_MAX_BRANCHES];  // float offset of a branch's net inside a member
    int col_off[ASAC_MAX_BRANCHES];         // first output column of a branch
    int S, H, depth, D;
};
__host__ __device__ __forceinline__ NetShape dnet_shape(const DNets &n, int k) {
    return NetShape{n.S, n.H, n.depth, n.sizes[k]};
}

struct DFwdArgs {
    DNets nets;
    const float *x;        // row r at x + r * x_row_stride
    int64_t x_row_stride;
    int rows;
    float *out;            // [members, rows, D]
};

struct DFwdPlan {
    int lda, wsz, off_a, off_b, off_ho, off_part, off_pipe, off_slots, n_slots, total;
};
__host__ __device__ __forceinline__ DFwdPlan dfwd_plan(int S, int H, int depth, int max_out) {
    DFwdPlan p;
    p.lda = tile_lda(H, S);
    p.wsz = round_up(tile_wsz(H, S), 256);
    int o = PASS_ROWS * p.lda;  // xin at 0
    p.off_a = o; o += PASS_ROWS * p.lda;
    p.off_b = o; o += PASS_ROWS * p.lda;
    p.off_ho = o; o += round_up(PASS_ROWS * max_out, 4);
    p.off_part = o; o += tile_part_floats(H);
    p.off_pipe = o; o += PIPE_HEADER_FLOATS;
    p.off_slots = o;
    o += 256;
    p.n_slots = slots_that_fit(o, p.wsz, depth);
    p.total = o + p.n_slots * p.wsz;
    return p;
}

// grid (row tiles, members * branches)
__global__ void __launch_bounds__(NT) k_dnets_forward(const DFwdArgs a) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const int tid = threadIdx.x;
    const DNets &n = a.nets;
    const int member = blockIdx.y / n.branches, k = blockIdx.y - member * n.branches;
    const NetShape s = dnet_shape(n, k);
    const float *prm = n.params + member * n.member_stride + n.branch_off[k];
    const DFwdPlan pl = dfwd_plan(n.S, n.H, n.depth, D_MAX_COLS);
    const int lda = pl.lda;
    const int r0 = blockIdx.x * PASS_ROWS;
    const int rows = min(PASS_ROWS, a.rows - r0);
    float *xin = sm, *bufA = sm + pl.off_a, *bufB = sm + pl.off_b, *ho = sm + pl.off_ho, *part = sm + pl.off_part;
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    if (tid == 0) push_trunk_jobs(jobs, 0, s, prm);
    __syncthreads();
    WeightPipe pipe;
    pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, s.depth);
    const int K4 = round_up(s.in_dim, 4);
    for (int i = tid; i < PASS_ROWS * K4; i += NT) {
        const int r = i / K4, col = i - r * K4;
        xin[r * lda + col] = (r < rows && col < s.in_dim) ? a.x[(int64_t)(r0 + r) * a.x_row_stride + col] : 0.f;
    }
    __syncthreads();
    float *h = net_trunk_forward(s, pipe, xin, bufA, bufB, nullptr, nullptr, lda, PASS_ROWS, part);
    head_forward(h, lda, s.hidden, prm + net_w_off(s, s.depth), prm + net_b_off(s, s.depth), s.out_dim, rows, ho);
    __syncthreads();
    float *out = a.out + ((int64_t)member * a.rows + r0) * n.D + n.col_off[k];
    for (int i = tid; i < rows * s.out_dim; i += NT) {
        const int r = i / s.out_dim, j = i - r * s.out_dim;
        out[(int64_t)r * n.D + j] = ho[i];
    }
}

struct DBwdArgs {
    DNets nets;
    const float *x;
    int64_t x_row_stride;
    int rows;
    const float *d_out;    // [members, rows, D]  d loss / d output (already scaled: mean over the batch etc.)
    float *grad_part;      // [row tiles, members, member_floats]  partial gradients in the nets' flat layout
    int64_t member_floats;
    float *d_x;            // optional [members, branches, rows, S]: d loss / d input row (trained representation)
};

struct DBwdPlan {
    int lda, wsz, off_px, off_pz, off_g0, off_g1, off_g2, off_dO, off_part, off_pipe, off_slots, n_slots, n_jobs, total;
};
__host__ __device__ __forceinline__ DBwdPlan dbwd_plan(int S, int H, int depth, int max_out) {
    DBwdPlan p;
    p.lda = tile_lda(H, S);
    p.wsz = round_up(tile_wsz(H, S), 256);
    const int R = PASS_ROWS;
    int o = 0;
    p.off_px = o; o += (depth + 1) * R * p.lda;
    p.off_pz = o; o += depth * R * p.lda;
    p.off_g0 = o; o += R * p.lda;
    p.off_g1 = o; o += R * p.lda;
    p.off_g2 = o; o += R * p.lda;
    p.off_dO = o; o += round_up(R * max_out, 4);
    p.off_part = o; o += tile_part_floats(H);
    p.off_pipe = o; o += PIPE_HEADER_FLOATS;
    p.off_slots = o;
    o += 256;
    p.n_jobs = 2 * depth - 1;
    p.n_slots = slots_that_fit(o, p.wsz, p.n_jobs);
    p.total = o + p.n_slots * p.wsz;
    return p;
}

// grid (row tiles, members * branches): forward with saved activations, head backward, ResBlocks in reverse
__global__ void __launch_bounds__(NT) k_dnets_backward(const DBwdArgs a) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const int tid = threadIdx.x;
    const DNets &n = a.nets;
    const int member = blockIdx.y / n.branches, k = blockIdx.y - member * n.branches;
    const NetShape s = dnet_shape(n, k);
    const int d = s.depth, H = s.hidden, O = s.out_dim, R = PASS_ROWS;
    const float *prm = n.params + member * n.member_stride + n.branch_off[k];
    float *gout = a.grad_part + ((int64_t)blockIdx.x * n.members + member) * a.member_floats + n.branch_off[k];
    const DBwdPlan pl = dbwd_plan(n.S, n.H, n.depth, D_MAX_COLS);
    const int lda = pl.lda;
    const int r0 = blockIdx.x * R;
    const int rows = min(R, a.rows - r0);
    const LayerBufs px(sm + pl.off_px, R * lda), pz(sm + pl.off_pz, R * lda);
    float *g[3] = {sm + pl.off_g0, sm + pl.off_g1, sm + pl.off_g2};
    float *dO = sm + pl.off_dO, *part = sm + pl.off_part;
    WeightJob *jobs = reinterpret_cast<WeightJob *>(sm + pl.off_pipe);
    uint64_t *bars = reinterpret_cast<uint64_t *>(jobs + MAX_WEIGHT_JOBS);
    if (tid == 0) {
        const int nj = push_trunk_jobs(jobs, 0, s, prm);
        push_trunk_jobs_reverse(jobs, nj, s, prm);
    }
    __syncthreads();
    WeightPipe pipe;
    pipe_init(pipe, aligned_slots(sm, pl.off_slots), bars, jobs, pl.n_slots, pl.wsz, pl.n_jobs);
    const int K4 = round_up(s.in_dim, 4);
    for (int i = tid; i < R * K4; i += NT) {
        const int r = i / K4, col = i - r * K4;
        px[0][r * lda + col] = (r < rows && col < s.in_dim) ? a.x[(int64_t)(r0 + r) * a.x_row_stride + col] : 0.f;
    }
    const float *d_out = a.d_out + ((int64_t)member * a.rows + r0) * n.D + n.col_off[k];
    for (int i = tid; i < R * O; i += NT) {
        const int r = i / O, j = i - r * O;
        dO[i] = r < rows ? d_out[(int64_t)r * n.D + j] : 0.f;
    }
    __syncthreads();
    net_trunk_forward(s, pipe, px[0], nullptr, nullptr, px, pz, lda, R, part);
    const float *Wh = prm + net_w_off(s, d);
    head_backward(dO, O, px[d], lda, H, Wh, R, gout + net_w_off(s, d), gout + net_b_off(s, d), g[0], lda);
    int cur = 0;
#pragma unroll 1
    for (int l = d - 1; l >= 0; --l) {
        const int K = net_k(s, l);
        __syncthreads();
        float *dY = g[cur], *dZ = g[(cur + 1) % 3], *dX = g[(cur + 2) % 3];
        gelu_backward(dY, pz[l], dZ, lda, H, rows);
        __syncthreads();
        layer_weight_grad(dZ, lda, px[l], lda, H, K, rows, gout + net_w_off(s, l), gout + net_b_off(s, l));
        if (l > 0) {
            const float *Ws, *bs;
            const bool swz = pipe_front_swizzled(pipe);
            pipe_acquire(pipe, Ws, bs);
            layer_input_grad(H, dZ, lda, Ws, dY, dX, true, part, swz);  // ends with a CTA barrier
            pipe_release(pipe);
            cur = (cur + 2) % 3;
        } else if (a.d_x) {
            // first layer's input gradient (the critics' share of d loss / d state, sac_base.py:1573-1601)
            const int S = s.in_dim;
            const float *W0 = prm + net_w_off(s, 0);
            float *dx = a.d_x + (((int64_t)member * n.branches + k) * a.rows + r0) * S;
            for (int t = tid; t < rows * S; t += NT) {
                const int r = t / S, kk = t - r * S;
                float acc = 0.f;
                for (int h = 0; h < H; ++h) acc = fmaf(dZ[r * lda + h], __ldg(W0 + (int64_t)h * S + kk), acc);
                if (S == H) acc += dY[r * lda + kk];  // residual first block
                dx[t] = acc;
            }
        }
    }
}

// ---------------------------------------------------------------- per-row math
// per-branch log-softmax of `logits` (D columns) -> logn (normalised logits), p (probabilities)
__device__ __forceinline__ void d_softmax(const float *logits, const int *sizes, int branches, float *logn, float *p) {
    int c = 0;
    for (int k = 0; k < branches; ++k) {
        float m = -INFINITY;
        for (int j = 0; j < sizes[k]; ++j) m = fmaxf(m, logits[c + j]);
        float se = 0.f;
        for (int j = 0; j < sizes[k]; ++j) se += expf(logits[c + j] - m);
        const float lse = m + logf(se);           // torch.logsumexp
        float m2 = -INFINITY;
        for (int j = 0; j < sizes[k]; ++j) { logn[c + j] = logits[c + j] - lse; m2 = fmaxf(m2, logn[c + j]); }
        float s2 = 0.f;
        for (int j = 0; j < sizes[k]; ++j) { p[c + j] = expf(logn[c + j] - m2); s2 += p[c + j]; }
        for (int j = 0; j < sizes[k]; ++j) p[c + j] = p[c + j] / s2;   // torch.softmax of the normalised logits
        c += sizes[k];
    }
}

struct DTargetArgs {
    AsacSacConfig cfg;
    int branches, D, AF;            // AF = D + A: width of the stored action / mu_prob rows
    int sizes[ASAC_MAX_BRANCHES];
    const float *pi_logits;         // [B * L, D] policy logits on every row of the window
    const float *tq;                // [E, B * L, D] target critics on every row
    const float *actions_full, *mu_full;   // [B, L, AF]
    const float *pi_probs_d;        // post: [B, L - 1, D] probabilities just computed (mu := pi, sac_base.py:2233)
    const float *rewards;
    const uint8_t *dones, *last_masks, *padding_masks;
    const float *log_d_alpha;
    int post;
    float *d_y;                     // [B]
};

// one thread per batch element (sac_base.py:1384-1412, 1244-1295)
__global__ void __launch_bounds__(256) k_d_target(const DTargetArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const AsacSacConfig &c = a.cfg;
    if (e >= c.batch) return;
    const int L = c.seq_len, b = c.burn_in, n = c.n_step, D = a.D, E = c.ensemble, K = a.branches;
    const float alpha = expf(a.log_d_alpha[0]);
    float logn[D_MAX_COLS], p[D_MAX_COLS];
    float v[ASAC_MAX_NSTEP + 1], ratio[ASAC_MAX_NSTEP];
    for (int k = 0; k <= n; ++k) {
        const int64_t row = (int64_t)e * L + b + k;
        d_softmax(a.pi_logits + row * D, a.sizes, K, logn, p);
        float s = 0.f;
        for (int j = 0; j < D; ++j) {
            float q = 0.f;
            for (int i = 0; i < E; ++i) q += a.tq[((int64_t)i * c.batch * L + row) * D + j];
            q = q / (float)E;                                    // torch.stack(...).mean(0)
            s += p[j] * (q - alpha * logf(fmaxf(p[j], 1e-8f)));
        }
        v[k] = s / (float)K;
        if (c.use_n_step_is && k < n) {
            const float *act = a.actions_full + row * a.AF;
            const float *mu = a.post ? a.pi_probs_d + ((int64_t)e * (L - 1) + b + k) * D : a.mu_full + row * a.AF;
            float lp = 0.f, mprod = 1.f;
            int cc = 0;
            for (int kk = 0; kk < K; ++kk) {
                int best = 0;
                for (int j = 1; j < a.sizes[kk]; ++j)
                    if (act[cc + j] > act[cc + best]) best = j;   // value.max(-1)[1]
                lp += logn[cc + best];
                cc += a.sizes[kk];
            }
            for (int j = 0; j < D; ++j) {
                const float m = mu[j] * act[j];
                mprod *= (m == 0.f) ? 1.f : m;
            }
            ratio[k] = expf(lp) / fmaxf(mprod, 1e-8f);            // sac_base.py:1275
        }
    }
    float sum = 0.f, cprod = 1.f;
    for (int k = 0; k < n; ++k) {
        const int64_t idx = (int64_t)e * c.bn_stride + b + k;
        const float nd = a.dones[idx] ? 0.f : 1.f;
        float td = a.rewards[idx] + (c.gamma * nd) * v[k + 1] - v[k];
        td = c.gamma_ratio[k] * td;
        if (c.use_n_step_is) {
            td = c.lambda_ratio[k] * td;
            const float rho = fminf(ratio[k], c.v_rho);
            td = (cprod * rho) * td;
            cprod = cprod * fminf(ratio[k], c.v_c);
        }
        const float keep = (a.last_masks[idx] | a.padding_masks[idx]) ? 0.f : 1.f;
        sum += td * keep;
    }
    a.d_y[e] = v[0] + sum;
}

struct DDqnArgs {
    AsacSacConfig cfg;
    int branches, D;
    int sizes[ASAC_MAX_BRANCHES];
    const float *eval_q;            // [E, B * L, D] ONLINE critics on every row of the window
    const float *tq;                // [E, B * L, D] target critics
    const int32_t *perm_target, *perm_online;   // [E] each (the reference's two randperm draws), NULL: identity
    int Es;                         // pairs taken: ensemble_q_sample
    const float *rewards;
    const uint8_t *dones, *last_masks, *padding_masks;
    float *d_y;
};
// get_dqn_like_d_y (sac_base.py:1193-1242, 1363-1383): double DQN on the last solid step — per pair i of the two
// shuffled stacks, the ONLINE member picks the arg-max action of every branch, the TARGET member is evaluated there;
// min over the pairs; n-step discounted rewards in front.  One thread per batch element.
__global__ void __launch_bounds__(256) k_d_target_dqn(const DDqnArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const AsacSacConfig &c = a.cfg;
    if (e >= c.batch) return;
    const int L = c.seq_len, b = c.burn_in, n = c.n_step, D = a.D, K = a.branches;
    int last = n - 1;  // get_last_false_indexes: n - 1 when every step is solid-masked (argmin of all ones is 0)
    for (int k = n - 1; k >= 0; --k) {
        const int64_t idx = (int64_t)e * c.bn_stride + b + k;
        if (!(a.last_masks[idx] | a.padding_masks[idx])) { last = k; break; }
    }
    const int64_t row = (int64_t)e * L + b + last + 1;
    float next_q = INFINITY;
    for (int j = 0; j < a.Es; ++j) {
        const int it = a.perm_target ? a.perm_target[j] : j, io = a.perm_online ? a.perm_online[j] : j;
        const float *qe = a.eval_q + ((int64_t)io * c.batch * L + row) * D;
        const float *qt = a.tq + ((int64_t)it * c.batch * L + row) * D;
        float s = 0.f;
        int cc = 0;
        for (int k = 0; k < K; ++k) {
            int best = 0;
            for (int jj = 1; jj < a.sizes[k]; ++jj)
                if (qe[cc + jj] > qe[cc + best]) best = jj;   // torch.argmax: first maximum
            s += qt[cc + best];
            cc += a.sizes[k];
        }
        next_q = fminf(next_q, s / (float)K);
    }
    float g = 0.f;
    for (int k = 0; k < n; ++k) g += c.gamma_ratio[k] * a.rewards[(int64_t)e * c.bn_stride + b + k];
    const float nd = a.dones[(int64_t)e * c.bn_stride + b + last] ? 0.f : 1.f;
    a.d_y[e] = g + powf(c.gamma, (float)(last + 1)) * next_q * nd;
}

struct DQGradArgs {
    int B, L, b, E, branches, D, AF;
    const float *q;            // [E, B, D] online critics on (s_b)
    const float *actions_full; // [B, L, AF]
    const float *d_y, *weights;
    float scale;               // 2 when the reference doubles the discrete loss (hybrid, clip_epsilon <= 0)
    float *d_out;              // [E, B, D]
    float *loss;               // [E, B]  weighted per-sample loss of the discrete part
    float *q_single;           // [E, B]
};
// sac_base.py:1543-1547, 1564-1570: loss_i = mean_B((sum(onehot * q_i) / branches - d_y)^2 * w)
__global__ void __launch_bounds__(256) k_d_q_grad(const DQGradArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.E * a.B) return;
    const int i = t / a.B, e = t - i * a.B;
    const float *q = a.q + (int64_t)t * a.D;
    const float *act = a.actions_full + ((int64_t)e * a.L + a.b) * a.AF;
    float qs = 0.f;
    for (int j = 0; j < a.D; ++j) qs += act[j] * q[j];
    qs = qs / (float)a.branches;
    const float w = a.weights ? a.weights[e] : 1.f;
    const float diff = qs - a.d_y[e];
    a.loss[t] = a.scale * diff * diff * w;
    a.q_single[t] = qs;
    const float gl = a.scale * 2.f * diff * w / (float)a.B;
    for (int j = 0; j < a.D; ++j) a.d_out[(int64_t)t * a.D + j] = gl * act[j] / (float)a.branches;
}

struct DPiGradArgs {
    int B, L, b, E, branches, D, AF, stride_rows;  // logits row of element e: (e * stride_rows + row_off)
    int row_off;
    int sizes[ASAC_MAX_BRANCHES];
    const float *logits;       // policy logits
    const float *q;            // [E, B, D] online critics (after their step) on s_b
    const float *mu_full;      // [B, L, AF]
    const float *log_d_alpha;
    float penalty;
    float *d_out;              // [B, D]
    float *loss, *entropy;     // [B] each
};
// sac_base.py:1858-1880
__global__ void __launch_bounds__(256) k_d_pi_grad(const DPiGradArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.B) return;
    const int D = a.D, K = a.branches;
    float logn[D_MAX_COLS], p[D_MAX_COLS], gq[D_MAX_COLS];
    d_softmax(a.logits + ((int64_t)e * a.stride_rows + a.row_off) * D, a.sizes, K, logn, p);
    const float alpha = expf(a.log_d_alpha[0]);
    float loss = 0.f;
    for (int j = 0; j < D; ++j) {
        float q = 0.f;
        for (int i = 0; i < a.E; ++i) q += a.q[((int64_t)i * a.B + e) * D + j];
        q = q / (float)a.E;
        const float inner = alpha * logf(fmaxf(p[j], 1e-8f)) - q;
        loss += p[j] * inner;
        gq[j] = (inner + (p[j] >= 1e-8f ? alpha : 0.f)) / (float)K;   // d(sum p * inner / K) / d p_j
    }
    loss = loss / (float)K;
    const float *mu = a.mu_full + ((int64_t)e * a.L + a.b) * a.AF;
    float mu_ent = 0.f;
    for (int j = 0; j < D; ++j) mu_ent -= mu[j] * logf(fmaxf(mu[j], 1e-8f));
    mu_ent = mu_ent / (float)K;
    float ent_k[ASAC_MAX_BRANCHES], pi_ent = 0.f;
    int c = 0;
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int j = 0; j < a.sizes[k]; ++j) s -= fmaxf(logn[c + j], -3.4028234663852886e38f) * p[c + j];
        ent_k[k] = s;
        pi_ent += s;
        c += a.sizes[k];
    }
    pi_ent = pi_ent / (float)K;
    const float gap = mu_ent - pi_ent;
    loss += a.penalty * (gap * gap / 2.f);
    a.loss[e] = loss;
    a.entropy[e] = pi_ent;
    // d loss / d logits through the per-branch softmax, / B for the mean over the batch
    c = 0;
    for (int k = 0; k < K; ++k) {
        float dot = 0.f;
        for (int j = 0; j < a.sizes[k]; ++j) dot += p[c + j] * gq[c + j];
        for (int j = 0; j < a.sizes[k]; ++j) {
            const float d1 = p[c + j] * (gq[c + j] - dot);
            const float dent = -p[c + j] * (logn[c + j] + ent_k[k]) / (float)K;   // d pi_ent / d logit
            a.d_out[(int64_t)e * D + c + j] = (d1 - a.penalty * gap * dent) / (float)a.B;
        }
        c += a.sizes[k];
    }
}

struct DPostArgs {
    int B, L, b, E, branches, D, AF;
    int sizes[ASAC_MAX_BRANCHES];
    const float *logits;       // [B * L, D] policy logits after the policy step
    float *pi_probs_d;         // [B, L - 1, D]
    float *pi_probs_full;      // [B, L - 1, AF] columns [0, D) are written (may be null)
};
// get_l_probs, discrete part (sac_base.py:1176-1178)
__global__ void __launch_bounds__(256) k_d_probs(const DPostArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.B * (a.L - 1)) return;
    const int e = t / (a.L - 1), r = t - e * (a.L - 1);
    float logn[D_MAX_COLS], p[D_MAX_COLS];
    d_softmax(a.logits + ((int64_t)e * a.L + r) * a.D, a.sizes, a.branches, logn, p);
    for (int j = 0; j < a.D; ++j) {
        a.pi_probs_d[(int64_t)t * a.D + j] = p[j];
        if (a.pi_probs_full) a.pi_probs_full[(int64_t)t * a.AF + j] = p[j];
    }
}

struct DAlphaArgs {
    int B, L, b, branches, D;
    int sizes[ASAC_MAX_BRANCHES];
    float target_ratio;        // target_d_alpha: the per-column target is ratio * log(size of the branch)
    const float *logits;       // [B * L, D] policy logits after the policy step
    float *log_d_alpha, *m, *v;
    const int64_t *step;       // Adam steps taken so far by the alpha optimizer
    double lr;
    float grad_scale;
    float *grad_out, *loss_out;
};
// _train_alpha, discrete part (sac_base.py:1924-1929) + torch.optim.Adam on log_d_alpha; one CTA
__global__ void __launch_bounds__(NT) k_d_alpha(const DAlphaArgs a) {  // NT threads: block_sum's contract
    __shared__ float red[32];
    float acc = 0.f;
    for (int e = threadIdx.x; e < a.B; e += blockDim.x) {
        float logn[D_MAX_COLS], p[D_MAX_COLS];
        d_softmax(a.logits + ((int64_t)e * a.L + a.b) * a.D, a.sizes, a.branches, logn, p);
        float s = 0.f;
        int c = 0;
        for (int k = 0; k < a.branches; ++k) {
            const float target = a.target_ratio * logf((float)a.sizes[k]);
            for (int j = 0; j < a.sizes[k]; ++j) s += p[c + j] * (-logf(fmaxf(p[c + j], 1e-8f)) - target);
            c += a.sizes[k];
        }
        acc += s / (float)a.branches;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const float g = (acc / (float)a.B) * a.grad_scale;
        a.grad_out[0] = g;
        a.loss_out[0] = a.log_d_alpha[0] * (acc / (float)a.B);
        const double t = (double)(a.step[0] + 1);
        const float m = 0.9f * a.m[0] + 0.1f * g;
        const float v = 0.999f * a.v[0] + 0.001f * (g * g);
        a.m[0] = m; a.v[0] = v;
        const double bc1 = 1.0 - beta_pow(ADAM_LN_BETA1, t), bc2 = 1.0 - beta_pow(ADAM_LN_BETA2, t);
        const float step_size = (float)(a.lr / bc1);
        const float denom = sqrtf(v) / (float)sqrt(bc2) + 1e-8f;
        a.log_d_alpha[0] = a.log_d_alpha[0] - step_size * (m / denom);
    }
}

struct DTdArgs {
    int B, L, b, E, branches, D, AF, accumulate;
    const float *q;            // [E, B, D] online critics on s_b
    const float *actions_full;
    const float *d_y;          // [B] y of the td-error pass
    float *td_error;           // [B]
};
// _get_td_error, discrete part (sac_base.py:2226-2230): mean_i |sum(onehot * q_i) / branches - d_y|
__global__ void __launch_bounds__(256) k_d_td(const DTdArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.B) return;
    const float *act = a.actions_full + ((int64_t)e * a.L + a.b) * a.AF;
    float s = 0.f;
    for (int i = 0; i < a.E; ++i) {
        float qs = 0.f;
        for (int j = 0; j < a.D; ++j) qs += act[j] * a.q[((int64_t)i * a.B + e) * a.D + j];
        s += fabsf(qs / (float)a.branches - a.d_y[e]);
    }
    s = s / (float)a.E;
    a.td_error[e] = a.accumulate ? a.td_error[e] + s : s;
}
