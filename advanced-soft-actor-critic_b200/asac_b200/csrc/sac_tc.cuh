// The SAC update kernels on the tcgen05 layer engine (tc_engine.cuh): every hidden layer and every head of
// the policy / critic MLPs is a batch of tcgen05.mma with the accumulator in tensor memory; CUDA cores do what
// is left — bias, exact-erf GELU, residual, the per-row transcendental stages, V-trace, losses.
// Included by sac.cu (uses its SacArgs, shapes and scalar helpers).  Stock width 64 only: UMMA_M is the hidden
// width; other widths stay on the FFMA row-tile kernels.
//
// Reference path replaced: ModelPolicy / ModelQ forwards inside _get_y, get_l_probs, _get_td_error,
// _train_rep_q and _train_policy (algorithm/sac_base.py:1159-1189, 1297-1466, 1516-1603, 1882-1908,
// 2182-2245; nn_models/layers/linear_layers.py:46-56).
// (no include guard games: included exactly once, INSIDE namespace asac, by sac.cu after its helpers)
#pragma once

// ---------------------------------------------------------------- value pass
struct ValueTcPlan {
    int RA;  // rows of the operand planes (multiple of 8)
    int off_xhi, off_xlo, off_w, off_bias, off_ho, off_qo, off_xs, off_logp, off_qmin, off_qpush, off_qmin2, off_ratio, off_qs,
        off_red, off_misc;
    int total;  // floats
};
__host__ __device__ __forceinline__ ValueTcPlan value_tc_plan(const AsacSacConfig &c, int TB, int mode) {
    ValueTcPlan p;
    const int L = c.seq_len, n = c.n_step, A = c.action_size;
    const int t0 = (mode == 1 && c.use_n_step_is) ? 0 : c.burn_in;
    const int Lp = L - t0;
    const int extra = (mode == 1 && c.rep_kind != 0) ? TB * (n + 1) : 0;
    const int rp = round_up(TB * Lp + extra, 8);
    const int rq = round_up(TB * (n + 1) + TB, 8);
    p.RA = rp > rq ? rp : rq;
    int o = 0;
    p.off_xhi = o; o += p.RA * TCF_M;
    p.off_xlo = o; o += p.RA * TCF_M;
    p.off_w = o; o += 4 * TCF_M * TCF_M;
    p.off_bias = o; o += 2 * TCF_M;
    p.off_ho = o; o += round_up(rp * 2 * A, 4);
    p.off_qo = o; o += round_up(rq, 4);
    p.off_xs = o; o += round_up(TB * (n + 1) * A, 4);
    p.off_logp = o; o += round_up(TB * (n + 1), 4);
    p.off_qmin = o; o += round_up(TB * (n + 1), 4);
    p.off_qpush = o; o += (c.ensemble - 1) * round_up(TB * (n + 1), 4);  // ranks 1.. push their value rows to rank 0
    p.off_qmin2 = o; o += ensemble_subset(c) ? round_up(TB * (n + 1), 4) : 0;
    p.off_ratio = o; o += round_up(TB * (n > 0 ? n : 1), 4);
    p.off_qs = o; o += round_up(c.ensemble * TB, 4);
    p.off_red = o; o += 32;
    p.off_misc = o; o += 8;
    p.total = o;
    return p;
}

// Same contract as k_value_pass (sac.cu): mode 0 = _get_y + target-Q(s_b, a_b), mode 1 = alpha-loss terms,
// get_l_probs, y' parts and Q_i(s_b, a_b).  grid (n_tiles, E), cluster (1, E, 1): every rank runs the policy and
// ONE ensemble member; rank 0 combines over distributed shared memory.
__global__ void __launch_bounds__(NT, 1) k_value_pass_tc(const __grid_constant__ SacArgs a) {
    // programmatic dependent launch: as in k_value_pass, the post pass stages its rows while the policy's Adam step
    // drains and waits in front of the first read of the policy
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
    const float *st_p = post && a.bat.states_post ? a.bat.states_post : a.bat.states;
    const float *st_v = post && a.bat.target_states ? a.bat.target_states : st_p;
    const bool split = post && c.rep_kind != 0;
    const int RPt = RP + (split ? RV : 0);
    const ValueTcPlan &pl = *reinterpret_cast<const ValueTcPlan *>(a.plan);  // computed on the host
    float *ho = sm + pl.off_ho, *qo = sm + pl.off_qo, *xs = sm + pl.off_xs, *logp = sm + pl.off_logp;
    float *qmin = sm + pl.off_qmin, *ratio = sm + pl.off_ratio, *qs = sm + pl.off_qs, *red = sm + pl.off_red;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + pl.off_misc + 2);

    const NetShape ps = pi_shape(c), qsh = q_shape(c);
    const int64_t q_stride = net_stride(qsh);
    const float *prm_pi = a.prm.pi, *prm_qt = a.prm.q_target + net * q_stride, *prm_q = a.prm.q + net * q_stride;

    TcfCtx cx;
    cx.x_hi = sm + pl.off_xhi; cx.x_lo = sm + pl.off_xlo;
    for (int q = 0; q < 2; ++q)
        for (int h = 0; h < 2; ++h) cx.w[q][h] = sm + pl.off_w + (2 * q + h) * TCF_M * TCF_M;
    cx.bias = sm + pl.off_bias;
    cx.bar = reinterpret_cast<uint64_t *>(sm + pl.off_misc);
    cx.phase = 0; cx.slot = 0;
    // cross terms of the 3xTF32 product in their own accumulator (see tcf_issue) whenever two accumulators fit
    // (measured on the stand-alone kernel, tools/tc_precision.py: a separate accumulator for the cross terms changes
    //  the error by < 1 % — the operand rounding in split_tf32 is what matters — so the pass keeps ONE accumulator
    //  and the smaller TMEM allocation, which lets CTAs of other kernels share the SM)
    const bool split_acc = false;
    cx.cross_cols = split_acc ? (uint32_t)pl.RA : 0u;
    const uint32_t tmem_cols = tcf_tmem_cols(pl.RA, split_acc);

    ASAC_PHASE(0, 0);
    if (tid < 32) tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(cx.bar, 1);
        fence_mbar_init();
    }
    // ---- the policy's first layer -> slot 0; P rows -> operand planes
    // Long windows (get_l_probs over a burn-in, sac_base.py:2563-2568): the policy rows are the bulk of the pass, so
    // the E cluster ranks each run a slice of them (whole 8-row groups) and hand their head outputs to the others
    // over distributed shared memory instead of all running every row.
    const int NP8 = round_up(RPt, 8) / 8;
    const bool share = E > 1 && NP8 >= 4 * E;
    const int p_lo = share ? (NP8 * net / E) * 8 : 0;
    const int p_hi = share ? (NP8 * (net + 1) / E) * 8 : NP8 * 8;
    TcfJob job = tcf_trunk_job(ps, prm_pi, 0);
    {
        TcfWeights w0;
        if (!late_wait) tcf_prefetch(w0, job);
        const int Kp = job.Kp;
        for (int i = tid; i < (p_hi - p_lo) * Kp; i += NT) {
            const int rl = i / Kp, col = i - rl * Kp, r = p_lo + rl;
            float v = 0.f;
            if (r < RP && col < S) {
                const int e = r / Lp, tt = r - e * Lp;
                v = st_p[((int64_t)(e0 + e) * L + t0 + tt) * S + col];
            } else if (r < RPt && col < S) {
                const int rv = r - RP, e = rv / (n + 1), k = rv - e * (n + 1);
                v = st_v[((int64_t)(e0 + e) * L + b + k) * S + col];
            }
            tcf_put(cx, rl, col, Kp, v);
        }
        if (late_wait) {
            pdl_wait();  // the policy's Adam step is complete and flushed from here on
            tcf_prefetch(w0, job);
        }
        tcf_store(w0, job, cx.w[0][0], cx.w[0][1], cx.bias);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cx.tmem = *tmem_slot;
    ASAC_PHASE(0, 1);

    // ---- policy over the rank's P rows
    {
        const int R = p_hi - p_lo;
        for (int l = 0; l < ps.depth; ++l) {
            const TcfJob nxt = l + 1 < ps.depth ? tcf_trunk_job(ps, prm_pi, l + 1) : tcf_head_job(ps, prm_pi);
            tcf_layer<false>(cx, job, &nxt, R, job.K == TCF_M, nullptr, 0, 0);
            job = nxt;
        }
        const TcfJob nxt = tcf_trunk_job(qsh, prm_qt, 0);
        tcf_layer<true>(cx, job, &nxt, R, false, ho + p_lo * 2 * A, 2 * A, min(RPt, p_hi) - p_lo);
        job = nxt;
        cluster.barrier_wait();  // (arrived at kernel entry: long complete)
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
    // ---- per P row: distribution, sampled action, log-probs, IS ratio, pi_probs, alpha terms  (as k_value_pass)
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
    // ---- critic inputs: V rows [state(e, b+k), tanh(x)], then (train, clipped loss) the S rows [state(e, b), a_b]
    const int K0 = S + A;
    auto stage_q_rows = [&](int first_v, int n_v, int n_s, const float *s_states) {
        const int Kp = round_up(K0, 8), Rp = round_up(n_v + n_s, 8);
        for (int i = tid; i < Rp * Kp; i += NT) {
            const int r = i / Kp, col = i - r * Kp;
            float v = 0.f;
            if (r < n_v) {
                const int rv = first_v + r, e = rv / (n + 1), k = rv - e * (n + 1);
                if (col < S) v = st_v[((int64_t)(e0 + e) * L + b + k) * S + col];
                else if (col < K0) v = tanhf(xs[rv * A + (col - S)]);
            } else if (r < n_v + n_s) {
                const int e = r - n_v;
                if (col < S) v = s_states[((int64_t)(e0 + e) * L + b) * S + col];
                else if (col < K0) v = a.bat.actions[((int64_t)(e0 + e) * c.bn_stride + b) * A + (col - S)];
            }
            tcf_put(cx, r, col, Kp, v);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        return Rp;
    };
    // ---- target critic `net` over the V rows (+ S rows for the clipped loss)
    {
        const int rq = need_tq ? RV + RS : RV;
        const int R = stage_q_rows(0, RV, need_tq ? RS : 0, st_p);
        for (int l = 0; l < qsh.depth; ++l) {
            const TcfJob nxt = l + 1 < qsh.depth ? tcf_trunk_job(qsh, prm_qt, l + 1) : tcf_head_job(qsh, prm_qt);
            tcf_layer<false>(cx, job, &nxt, R, job.K == TCF_M, nullptr, 0, 0);
            job = nxt;
        }
        const TcfJob nxt = tcf_trunk_job(qsh, prm_q, 0);
        tcf_layer<true>(cx, job, post ? &nxt : nullptr, R, false, qo, 1, rq);
        job = nxt;
        for (int r = tid; r < rq; r += NT) {
            if (r < RV) qmin[r] = qo[r];
            else a.wrk.tq[(int64_t)net * B + e0 + (r - RV)] = qo[r];
        }
        __syncthreads();
    }
    ASAC_PHASE(0, 5);
    // ---- post: online critic `net` over the S rows (sac_base.py:2211-2216)
    if (post) {
        const int R = stage_q_rows(0, 0, RS, st_p);
        for (int l = 0; l < qsh.depth; ++l) {
            const TcfJob nxt = l + 1 < qsh.depth ? tcf_trunk_job(qsh, prm_q, l + 1) : tcf_head_job(qsh, prm_q);
            tcf_layer<false>(cx, job, &nxt, R, job.K == TCF_M, nullptr, 0, 0);
            job = nxt;
        }
        tcf_layer<true>(cx, job, nullptr, R, false, qs, 1, RS);
    }
    ASAC_PHASE(0, 6);
    // ---- ensemble combine on rank 0 over distributed shared memory, in member order
    float *qmin2 = ensemble_subset(c) && a.bat.ensemble_perms ? sm + pl.off_qmin2 : qmin;
    if (!post && qmin2 == qmin && a.push_combine) {  // train pass, all members: push (see k_value_pass)
        const int stride = round_up(TB * (n + 1), 4);
        if (net != 0) {
            float *dst = cluster.map_shared_rank(sm + pl.off_qpush, 0) + (net - 1) * stride;
            for (int r = tid; r < RV; r += NT) dst[r] = qmin[r];
        }
        cluster.sync();
        tc_fence_after();
        if (tid < 32) tmem_dealloc(cx.tmem, tmem_cols);
        if (net != 0) return;
        for (int i = 1; i < E; ++i) {
            const float *src = sm + pl.off_qpush + (i - 1) * stride;
            for (int r = tid; r < RV; r += NT) qmin[r] = fminf(qmin[r], src[r]);
        }
        __syncthreads();
    } else {
    cluster.sync();
    if (net == 0) {
        combine_value_rows(cluster, c, a.bat.ensemble_perms, post ? 3 : 0, qmin, qmin2, RV);
        if (post)
            for (int i = 1; i < E; ++i) {
                const float *rqs = cluster.map_shared_rank(qs, i);
                for (int e = tid; e < TBa; e += NT) qs[i * TB + e] = rqs[e];
            }
    }
    cluster.sync();  // remote shared memory stays alive until rank 0 has read it
    ASAC_PHASE(0, 7);
    tc_fence_after();
    if (tid < 32) tmem_dealloc(cx.tmem, tmem_cols);
    if (net != 0) return;
    }

    // ---- per batch element: V, v-trace, y (sac_base.py:1244-1295, 1444-1464) — see k_value_pass
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
