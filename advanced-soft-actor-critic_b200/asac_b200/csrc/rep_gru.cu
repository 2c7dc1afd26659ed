// Recurrent representation kernels (sm_100a): the stock multi-layer GRU wrapper of the reference
// (algorithm/nn_models/layers/seq_layers.py:14-114, no padding mask) as used by a plugin ModelRep of
// the form of envs/test/nn_rnn.py:6-21 — forward over whole sampled windows and back-propagation
// through time of the critic loss's state gradient (sac_base.py:1510-1601, 2066-2105).
//
// The recurrence is a chain of dependent cells of a few hundred FMAs each: there is no parallelism
// inside a sequence worth a block barrier, so ONE WARP owns one sequence and the batch spreads over
// the SMs.  A lane owns a hidden unit (all three gates of it, so no gate exchange); when
// layers x width <= 32 the layers run as a WAVEFRONT — lane (l, j), layer l one step behind layer
// l - 1 — which turns L x layers dependent cells into L + layers - 1 stages, and for widths 8 / 16 / 32
// the lane's weight rows (forward) or columns (backward) live in registers.  Otherwise weights are
// read from shared memory, staged in rows of odd stride [W_ih row | W_hh row | b_ih | b_hh]
// (conflict-free for row-per-lane and column-per-lane reads).  Inputs, saved gates and hidden
// states of a sequence are staged once, with 8 loads per thread in flight, before the chain starts.
// Measured history (config 4, B200): forward 95 -> 22.5 us, backward 113 -> 34.8 us.
#include <math.h>

#include "common.cuh"

namespace asac {

__host__ __device__ __forceinline__ int gru_in(const AsacGruShape &s, int l) {
    return l == 0 ? s.obs_size + s.action_size : s.hidden;
}
__host__ __device__ __forceinline__ int64_t gru_layer_count(const AsacGruShape &s, int l) {
    return (int64_t)3 * s.hidden * (gru_in(s, l) + s.hidden + 2);
}
__host__ __device__ __forceinline__ int64_t gru_layer_off(const AsacGruShape &s, int l) {
    int64_t o = 0;
    for (int i = 0; i < l; ++i) o += gru_layer_count(s, i);
    return o;
}
__host__ __device__ __forceinline__ int64_t gru_count(const AsacGruShape &s) { return gru_layer_off(s, s.layers); }
// shared-memory row: in_l + H weights + 2 biases, odd stride
__host__ __device__ __forceinline__ int gru_row_stride(const AsacGruShape &s, int l) {
    return (gru_in(s, l) + s.hidden + 2) | 1;
}
__host__ __device__ __forceinline__ int gru_weight_floats(const AsacGruShape &s) {
    int o = 0;
    for (int l = 0; l < s.layers; ++l) o += 3 * s.hidden * gru_row_stride(s, l);
    return (o + 3) & ~3;
}

// i / d for 0 <= i < 2^20 and a divisor known only at run time, through its float reciprocal (three
// instructions instead of the ~25 of an integer division; (i + 0.5) / d is never within rounding of an integer)
struct FastDiv {
    int d;
    float inv;
    __device__ __forceinline__ explicit FastDiv(int d_) : d(d_), inv(1.f / (float)d_) {}
    __device__ __forceinline__ int div(int i) const { return __float2int_rz(((float)i + 0.5f) * inv); }
};

// Cooperative global -> shared staging with U loads of every thread in flight before the first store
// (the recurrent kernels start with a few thousand scattered floats; one load per loop trip would
// serialise that many DRAM / L2 round trips): element i comes from load(i) and goes to dst[where(i)].
template <int U, typename Load, typename Where>
__device__ __forceinline__ void staged_fill(float *dst, int n, Load load, Where where) {
    const int nt = blockDim.x;
    for (int base = threadIdx.x; base < n; base += U * nt) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * nt;
            v[u] = i < n ? load(i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * nt;
            if (i < n) dst[where(i)] = v[u];
        }
    }
}

// flat parameters of every layer -> padded shared-memory rows (whole CTA)
__device__ void gru_stage_weights(float *w_sm, const float *params, const AsacGruShape &s) {
    const int H = s.hidden;
    int base = 0;
    for (int l = 0; l < s.layers; ++l) {
        const int in = gru_in(s, l), rs = gru_row_stride(s, l);
        const float *p = params + gru_layer_off(s, l);
        const int n_ih = 3 * H * in, n_hh = 3 * H * H;
        const FastDiv by_in(in), by_h(H);
        staged_fill<8>(w_sm + base, n_ih + n_hh + 6 * H, [&](int i) { return __ldg(p + i); },
                       [&](int i) {
                           int g, c;
                           if (i < n_ih) { g = by_in.div(i); c = i - g * in; }
                           else if (i < n_ih + n_hh) { const int j = i - n_ih; g = by_h.div(j); c = in + (j - g * H); }
                           else if (i < n_ih + n_hh + 3 * H) { g = i - n_ih - n_hh; c = in + H; }
                           else { g = i - n_ih - n_hh - 3 * H; c = in + H + 1; }
                           return g * rs + c;
                       });
        base += 3 * H * rs;
    }
}

// x_t = [obs[b, t], pre_action[b, t]] for t < T of the CTA's n_seq sequences (seq0 ...) into
// region[w * region_stride + t * row_stride + c] (whole CTA); columns in0 .. row_stride - 1 are zeroed
__device__ __forceinline__ void gru_stage_inputs(float *region, int region_stride, int row_stride, const float *obs,
                                                 const float *actions, int bn_stride, const float *pre_actions,
                                                 int64_t seq0, int n_seq, int L, int T, int So, int A) {
    const int in0 = So + A, per = T * row_stride;
    const FastDiv by_per(per), by_row(row_stride);
    staged_fill<8>(region, n_seq * per,
                   [&](int i) {
                       const int w = by_per.div(i), r = i - w * per, t = by_row.div(r), c = r - t * row_stride;
                       const int64_t seq = seq0 + w;
                       if (c >= in0) return 0.f;
                       if (c < So) return obs[(seq * L + t) * So + c];
                       if (pre_actions) return pre_actions[(seq * L + t) * A + (c - So)];
                       return t > 0 ? actions[(seq * bn_stride + t - 1) * A + (c - So)] : 0.f;  // operators.py:39-59
                   },
                   [&](int i) { const int w = by_per.div(i); return w * region_stride + (i - w * per); });
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct GruFwdArgs {
    AsacGruShape s;
    AsacGruNet net[2];
    const float *obs, *actions, *pre_actions, *h0;
    int64_t h0_b_stride;
    int bn_stride, batch, seq_len;
};

constexpr int GRU_FWD_WARPS = 2;

// per-warp shared memory of the forward kernel: inputs xs[L][in0 padded to 4], layer 0's input
// projection gi0[L][3H] (used unless the projection runs inside the stage loop), hidden states h[layers][H]
__host__ __device__ __forceinline__ int gru_fwd_warp_floats(const AsacGruShape &s, int L) {
    return L * ((s.obs_size + s.action_size + 3) & ~3) + ((L * 3 * s.hidden + 3) & ~3) +
           ((s.layers * s.hidden + 3) & ~3);
}

// One GRU cell for hidden unit j of a layer whose three gate rows start at wr / wz / wn (shared
// memory, [W_ih row | W_hh row | b_ih | b_hh]).  gi != nullptr: the input projection (bias included)
// was computed up front (layer 0); otherwise xin is the layer below's output of this step.
struct GruCellOut {
    float r, z, n, ghn, h;
};
__device__ __forceinline__ GruCellOut gru_cell(const float *wr, const float *wz, const float *wn, int in, int H,
                                               const float *gi, const float *xin, const float *hp, int j) {
    float ir, iz, inn;
    if (gi) {
        ir = gi[j]; iz = gi[H + j]; inn = gi[2 * H + j];
    } else {
        ir = wr[in + H]; iz = wz[in + H]; inn = wn[in + H];
#pragma unroll 4
        for (int k = 0; k < in; ++k) {
            const float xv = xin[k];
            ir = fmaf(wr[k], xv, ir); iz = fmaf(wz[k], xv, iz); inn = fmaf(wn[k], xv, inn);
        }
    }
    float hr = wr[in + H + 1], hz = wz[in + H + 1], hn = wn[in + H + 1];
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        const float hv = hp[k];
        hr = fmaf(wr[in + k], hv, hr); hz = fmaf(wz[in + k], hv, hz); hn = fmaf(wn[in + k], hv, hn);
    }
    GruCellOut o;
    o.r = sigmoidf_(hr + ir);
    o.z = sigmoidf_(hz + iz);
    o.ghn = hn;
    o.n = tanhf(inn + hn * o.r);
    o.h = (hp[j] - o.n) * o.z + o.n;
    return o;
}

// Wavefront forward with the lane's weights in REGISTERS (widths 8 / 16 / 32): lane (l, j) owns unit j of
// layer l and keeps its three W_hh rows (and, above layer 0, its three W_ih rows) in registers; at stage
// st layer l runs step st - l.  Per stage a lane issues 2 x H/4 vector loads of the hidden vectors,
// 6H FMAs in six independent chains and the three activations: no weight traffic, no gate exchange,
// L + layers - 1 stages instead of L x layers dependent cells.  Lanes of layer 0 take their input
// projection from gi0 (computed for every step up front) and multiply the layer-below vector by
// zeros, so the warp does not diverge.
template <int H, bool MULTI>
__device__ __forceinline__ void gru_wave_forward(const AsacGruShape &s, const AsacGruNet &net, const float *w_sm,
                                                 const float *gi0, const float *xs, int in0p, bool inloop, float *hbuf,
                                                 int64_t seq, int L, int lane) {
    const int NL = s.layers;
    const int l = lane / H, j = lane - l * H;
    const bool valid = lane < NL * H;
    const int lc = valid ? l : 0;
    int wbase = 0;
    for (int i = 0; i < lc; ++i) wbase += 3 * H * gru_row_stride(s, i);
    const int in = gru_in(s, lc), rs = gru_row_stride(s, lc);
    const float *wr = w_sm + wbase + j * rs, *wz = wr + H * rs, *wn = wz + H * rs;
    float whh[3][H], wih[MULTI ? 3 : 1][MULTI ? H : 1];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        whh[0][k] = wr[in + k]; whh[1][k] = wz[in + k]; whh[2][k] = wn[in + k];
        if (MULTI) {
            // layer 0 (inloop: its inputs fit in H columns): W_ih rows zero-padded to H; else zeros (gi0 is used)
            const bool has = lc > 0 ? true : (inloop && k < in);
            wih[0][k] = has ? wr[k] : 0.f; wih[1][k] = has ? wz[k] : 0.f; wih[2][k] = has ? wn[k] : 0.f;
        }
    }
    const float bir = wr[in + H], biz = wz[in + H], bin = wn[in + H];
    const float bhr = wr[in + H + 1], bhz = wz[in + H + 1], bhn = wn[in + H + 1];
    const float4 *hp4 = reinterpret_cast<const float4 *>(hbuf + lc * H);
    const float4 *xp4 = reinterpret_cast<const float4 *>(hbuf + (lc > 0 ? lc - 1 : 0) * H);
    float *hself = hbuf + lc * H + j;
    // layer-0 lanes with inloop read x_t (row stride in0p <= H, zero padded) where the others read the layer below
    const int nx4 = (MULTI && lc == 0 && inloop) ? in0p / 4 : H / 4;
    // output cursors of this lane (advanced once per active stage: step t = 0, 1, ...)
    float *p_hn = net.hn ? net.hn + ((seq * L) * NL + lc) * H + j : nullptr;
    float *p_sv = net.save ? net.save + ((seq * L) * NL + lc) * 4 * H + j : nullptr;
    float *p_st = (valid && l == NL - 1) ? net.states + (seq * L) * H + j : nullptr;
#pragma unroll 1
    for (int st = 0; st < L + NL - 1; ++st) {
        const int t = st - l;
        const bool active = valid && t >= 0 && t < L;
        const int tc = min(max(t, 0), L - 1);
        const float *gi = gi0 + tc * 3 * H;
        const bool pre = lc == 0 && !inloop;  // input projection taken from the up-front pass
        float ir = pre ? gi[j] : bir, iz = pre ? gi[H + j] : biz, inn = pre ? gi[2 * H + j] : bin;
        if (MULTI && lc == 0) xp4 = reinterpret_cast<const float4 *>(inloop ? xs + tc * in0p : hbuf);
        float hr = bhr, hz = bhz, hn = bhn;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 hv = hp4[k4];
            const float h_[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                hr = fmaf(whh[0][k4 * 4 + c], h_[c], hr);
                hz = fmaf(whh[1][k4 * 4 + c], h_[c], hz);
                hn = fmaf(whh[2][k4 * 4 + c], h_[c], hn);
            }
            if (MULTI) {
                const float4 xv = k4 < nx4 ? xp4[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float x_[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    ir = fmaf(wih[0][k4 * 4 + c], x_[c], ir);
                    iz = fmaf(wih[1][k4 * 4 + c], x_[c], iz);
                    inn = fmaf(wih[2][k4 * 4 + c], x_[c], inn);
                }
            }
        }
        const float r = __frcp_rn(1.f + expf(-(hr + ir))), z = __frcp_rn(1.f + expf(-(hz + iz)));
        const float n = tanhf(inn + hn * r);
        const float hnew = (*hself - n) * z + n;
        __syncwarp();
        if (active) {
            *hself = hnew;
            if (p_hn) { *p_hn = hnew; p_hn += NL * H; }
            if (p_sv) {
                p_sv[0] = r; p_sv[H] = z; p_sv[2 * H] = n; p_sv[3 * H] = hn;
                p_sv += NL * 4 * H;
            }
            if (p_st) { *p_st = hnew; p_st += H; }
        }
        __syncwarp();
    }
}

// grid (ceil(B / GRU_FWD_WARPS), n_nets)
__global__ void __launch_bounds__(GRU_FWD_WARPS * 32) k_gru_forward(const __grid_constant__ GruFwdArgs a) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const AsacGruShape &s = a.s;
    const int H = s.hidden, NL = s.layers, L = a.seq_len, in0 = s.obs_size + s.action_size;
    const AsacGruNet &net = a.net[blockIdx.y];
    float *w_sm = sm;
    gru_stage_weights(w_sm, net.params, s);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t seq = (int64_t)blockIdx.x * GRU_FWD_WARPS + wid;
    float *xs = sm + gru_weight_floats(s) + wid * gru_fwd_warp_floats(s, L);
    const int in0p = (in0 + 3) & ~3;  // row stride of the staged inputs (zero padded)
    float *gi0 = xs + L * in0p, *hbuf = gi0 + ((L * 3 * H + 3) & ~3);  // 16-byte aligned regions
    // widths 8 / 16 with inputs no wider than the GRU: layer 0's lanes multiply x_t inside the stage loop
    const bool inloop = NL * H <= 32 && (H == 8 || H == 16) && in0p <= H;
    {
        const int64_t seq0 = (int64_t)blockIdx.x * GRU_FWD_WARPS;
        const int n_seq = (int)min((int64_t)GRU_FWD_WARPS, a.batch - seq0), wf = gru_fwd_warp_floats(s, L);
        float *first = sm + gru_weight_floats(s);
        gru_stage_inputs(first, wf, in0p, a.obs, a.actions, a.bn_stride, a.pre_actions, seq0, n_seq, L, L, s.obs_size,
                         s.action_size);
        const int hoff = (int)(hbuf - xs), nh = NL * H;
        const FastDiv by_nh(nh);
        staged_fill<2>(first + hoff, n_seq * nh,
                       [&](int i) { const int w = by_nh.div(i); return a.h0 ? a.h0[(seq0 + w) * a.h0_b_stride + (i - w * nh)] : 0.f; },
                       [&](int i) { const int w = by_nh.div(i); return w * wf + (i - w * nh); });
    }
    __syncthreads();
    if (seq >= a.batch) return;
    // layer 0's input projection W_ih x_t + b_ih of EVERY step, off the recurrent chain
    if (!inloop) {
        const int rs0 = gru_row_stride(s, 0);
        for (int idx = lane; idx < L * 3 * H; idx += 32) {
            const int t = idx / (3 * H), g = idx - t * 3 * H;
            const float *wr = w_sm + g * rs0, *x = xs + t * in0p;
            float acc = wr[in0 + H];
#pragma unroll 4
            for (int k = 0; k < in0; ++k) acc = fmaf(wr[k], x[k], acc);
            gi0[idx] = acc;
        }
    }
    __syncwarp();
    if (NL * H <= 32 && (H == 8 || H == 16 || H == 32)) {
        if (H == 8) gru_wave_forward<8, true>(s, net, w_sm, gi0, xs, in0p, inloop, hbuf, seq, L, lane);
        else if (H == 16) gru_wave_forward<16, true>(s, net, w_sm, gi0, xs, in0p, inloop, hbuf, seq, L, lane);
        else gru_wave_forward<32, false>(s, net, w_sm, gi0, xs, in0p, false, hbuf, seq, L, lane);
        return;
    }
    if (NL * H <= 32) {
        // wavefront: lane (l, j) owns unit j of layer l; at stage s layer l runs step s - l, reading the
        // layer below's output of that step (written one stage earlier) -> L + layers - 1 stages
        // instead of L x layers dependent cells, no gate exchange between lanes
        const int l = lane / H, j = lane - l * H;
        const bool valid = lane < NL * H;
        int wbase = 0;
        for (int i = 0; i < l && valid; ++i) wbase += 3 * H * gru_row_stride(s, i);
        const int in = valid ? gru_in(s, l) : 0, rs = valid ? gru_row_stride(s, l) : 0;
        const float *wr = w_sm + wbase + j * rs, *wz = wr + H * rs, *wn = wz + H * rs;
#pragma unroll 1
        for (int st = 0; st < L + NL - 1; ++st) {
            const int t = st - l;
            const bool active = valid && t >= 0 && t < L;
            GruCellOut o;
            if (active)
                o = gru_cell(wr, wz, wn, in, H, l == 0 ? gi0 + t * 3 * H : nullptr, hbuf + (l - 1) * H, hbuf + l * H, j);
            __syncwarp();
            if (active) {
                hbuf[l * H + j] = o.h;
                const int64_t cell = (seq * L + t) * NL + l;
                if (net.hn) net.hn[cell * H + j] = o.h;
                if (net.save) {
                    float *sv = net.save + cell * 4 * H;
                    sv[j] = o.r; sv[H + j] = o.z; sv[2 * H + j] = o.n; sv[3 * H + j] = o.ghn;
                }
                if (l == NL - 1) net.states[(seq * L + t) * H + j] = o.h;
            }
            __syncwarp();
        }
        return;
    }
    // wide / deep GRUs: layers one after the other, a lane owns units lane, lane + 32
#pragma unroll 1
    for (int t = 0; t < L; ++t) {
        int wbase = 0;
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            const int in = gru_in(s, l), rs = gru_row_stride(s, l);
            GruCellOut o[2];
            for (int u = 0, j = lane; j < H; ++u, j += 32) {
                const float *wr = w_sm + wbase + j * rs, *wz = wr + H * rs, *wn = wz + H * rs;
                o[u] = gru_cell(wr, wz, wn, in, H, l == 0 ? gi0 + t * 3 * H : nullptr, hbuf + (l - 1) * H, hbuf + l * H, j);
            }
            __syncwarp();
            for (int u = 0, j = lane; j < H; ++u, j += 32) {
                hbuf[l * H + j] = o[u].h;
                const int64_t cell = (seq * L + t) * NL + l;
                if (net.hn) net.hn[cell * H + j] = o[u].h;
                if (net.save) {
                    float *sv = net.save + cell * 4 * H;
                    sv[j] = o[u].r; sv[H + j] = o[u].z; sv[2 * H + j] = o[u].n; sv[3 * H + j] = o[u].ghn;
                }
                if (l == NL - 1) net.states[(seq * L + t) * H + j] = o[u].h;
            }
            __syncwarp();
            wbase += 3 * H * rs;
        }
    }
}

// ------------------------------------------------------------------------------------ backward
struct GruBwdArgs {
    AsacGruShape s;
    const float *params, *obs, *actions, *pre_actions, *h0, *grad_state, *hn, *save;
    float *grad_part;
    int64_t h0_b_stride, part_stride;
    int bn_stride, batch, seq_len, t_grad, ensemble, tile;
};
constexpr int GRU_BWD_THREADS = 256;

// per-sequence shared memory of the backward kernel (floats); T1 = t_grad + 1 steps carry gradient
struct GruBwdPlan {
    int off_xs, off_hn, off_h0, off_sv, off_dg, off_dh, off_dhd, off_dx, total;
};
__host__ __device__ __forceinline__ GruBwdPlan gru_bwd_plan(const AsacGruShape &s, int T1) {
    GruBwdPlan p;
    const int H = s.hidden, NL = s.layers;
    int o = 0;
    auto take = [&o](int n) { const int at = o; o += (n + 3) & ~3; return at; };  // 16-byte aligned regions
    p.off_xs = take(T1 * (s.obs_size + s.action_size));
    p.off_hn = take(T1 * NL * H);
    p.off_h0 = take(NL * H);
    p.off_sv = take(T1 * NL * 4 * H);
    p.off_dg = take(T1 * NL * 4 * H);
    p.off_dh = take(NL * H);
    p.off_dhd = take(NL * H);
    p.off_dx = take(NL * H);
    p.total = o;
    return p;
}

// Wavefront BPTT with the lane's weight COLUMNS in registers (see gru_wave_forward): lane (l, j) keeps
// W_hh[:, j] and (above layer 0) W_ih[:, j]; per stage it turns its unit's output gradient into the
// four gate gradients (shared memory, kept for the weight-gradient phase), then reads the layer's
// 4H gate gradients back as vectors and forms d h_{t-1}[j] and the input gradient for the layer below.
template <int H, bool MULTI>
__device__ __forceinline__ void gru_wave_backward(const AsacGruShape &s, const float *w_sm, float *me,
                                                  const GruBwdPlan &pl, int T1, int lane) {
    const int NL = s.layers;
    float *dh = me + pl.off_dh, *dhd = me + pl.off_dhd, *dx = me + pl.off_dx;
    const int l = lane / H, j = lane - l * H;
    const bool valid = lane < NL * H;
    const int lc = valid ? l : 0;
    int wbase = 0;
    for (int i = 0; i < lc; ++i) wbase += 3 * H * gru_row_stride(s, i);
    const int in = gru_in(s, lc), rs = gru_row_stride(s, lc);
    const float *wcol = w_sm + wbase;
    float chh[3 * H], cih[MULTI ? 3 * H : 1];
#pragma unroll
    for (int g = 0; g < 3 * H; ++g) {
        chh[g] = wcol[g * rs + in + j];
        if (MULTI) cih[g] = lc > 0 ? wcol[g * rs + j] : 0.f;
    }
#pragma unroll 1
    for (int st = 0; st < T1 + NL - 1; ++st) {
        const int t = (T1 - 1) - (st - (NL - 1 - l));
        const bool active = valid && t >= 0 && t <= T1 - 1;
        const int tc = min(max(t, 0), T1 - 1);
        float *dg = me + pl.off_dg + (tc * NL + lc) * 4 * H;
        if (active) {
            const float *sv = me + pl.off_sv + (t * NL + l) * 4 * H;
            const float hprev = t > 0 ? me[pl.off_hn + ((t - 1) * NL + l) * H + j] : me[pl.off_h0 + l * H + j];
            const float r = sv[j], z = sv[H + j], n = sv[2 * H + j], ghn = sv[3 * H + j];
            const float d = dh[l * H + j] + dx[l * H + j];
            const float dz = d * (hprev - n), dn = d * (1.f - z);
            const float dan = dn * (1.f - n * n);
            const float dr = dan * ghn;
            dg[j] = dr * (r * (1.f - r));
            dg[H + j] = dz * (z * (1.f - z));
            dg[2 * H + j] = dan;
            dg[3 * H + j] = dan * r;
            dhd[l * H + j] = d * z;
        }
        __syncwarp();
        float sh = dhd[lc * H + j], sx = 0.f;
        const float4 *dg4 = reinterpret_cast<const float4 *>(dg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {        // gate blocks: d a_r, d a_z, d (W_in x), d (W_hn h)
#pragma unroll
            for (int k4 = 0; k4 < H / 4; ++k4) {
                const float4 v = dg4[q * (H / 4) + k4];
                const float v_[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int u = k4 * 4 + c;
                    if (q < 2) sh = fmaf(chh[q * H + u], v_[c], sh);
                    if (q == 3) sh = fmaf(chh[2 * H + u], v_[c], sh);
                    if (MULTI && q < 3) sx = fmaf(cih[q * H + u], v_[c], sx);
                }
            }
        }
        __syncwarp();
        if (active) {
            dh[l * H + j] = sh;
            if (MULTI && l > 0) dx[(l - 1) * H + j] = sx;
        }
        __syncwarp();
    }
}

// grid ceil(B / tile); warp w < tile runs the BPTT of sequence blockIdx.x * tile + w, then the whole CTA
// turns the stored gate gradients into one partial weight gradient per sequence.
__global__ void __launch_bounds__(GRU_BWD_THREADS) k_gru_backward(const __grid_constant__ GruBwdArgs a) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    const AsacGruShape &s = a.s;
    const int H = s.hidden, NL = s.layers, L = a.seq_len, T1 = a.t_grad + 1, in0 = s.obs_size + s.action_size;
    const GruBwdPlan pl = gru_bwd_plan(s, T1);
    float *w_sm = sm;
    float *seq_sm = sm + gru_weight_floats(s);
    gru_stage_weights(w_sm, a.params, s);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t seq0 = (int64_t)blockIdx.x * a.tile;
    const int n_seq = (int)min((int64_t)a.tile, a.batch - seq0);
    {   // the tile's sequences: inputs, hidden states and gates of the first t_grad + 1 steps (whole CTA)
        gru_stage_inputs(seq_sm + pl.off_xs, pl.total, in0, a.obs, a.actions, a.bn_stride, a.pre_actions, seq0, n_seq, L, T1,
                         s.obs_size, s.action_size);
        const int n_hn = T1 * NL * H, n_sv = T1 * NL * 4 * H, n_h0 = NL * H;
        const FastDiv by_hn(n_hn), by_sv(n_sv), by_h0(n_h0);
        staged_fill<8>(seq_sm + pl.off_hn, n_seq * n_hn,
                       [&](int i) { const int w = by_hn.div(i); return a.hn[(seq0 + w) * L * NL * H + (i - w * n_hn)]; },
                       [&](int i) { const int w = by_hn.div(i); return w * pl.total + (i - w * n_hn); });
        staged_fill<8>(seq_sm + pl.off_sv, n_seq * n_sv,
                       [&](int i) { const int w = by_sv.div(i); return a.save[(seq0 + w) * L * NL * 4 * H + (i - w * n_sv)]; },
                       [&](int i) { const int w = by_sv.div(i); return w * pl.total + (i - w * n_sv); });
        staged_fill<2>(seq_sm + pl.off_h0, n_seq * n_h0,
                       [&](int i) { const int w = by_h0.div(i); return a.h0 ? a.h0[(seq0 + w) * a.h0_b_stride + (i - w * n_h0)] : 0.f; },
                       [&](int i) { const int w = by_h0.div(i); return w * pl.total + (i - w * n_h0); });
    }
    if (wid < n_seq) {
        const int64_t seq = seq0 + wid;
        float *me = seq_sm + wid * pl.total;
        for (int i = lane; i < NL * H; i += 32) {
            float g = 0.f;
            if (i >= (NL - 1) * H)  // d loss / d state[:, t_grad], summed over the critics in member order
                for (int e = 0; e < a.ensemble; ++e) g += a.grad_state[((int64_t)e * a.batch + seq) * H + (i - (NL - 1) * H)];
            me[pl.off_dh + i] = g;
            me[pl.off_dx + i] = 0.f;
        }
    }
    __syncthreads();
    if (wid < n_seq && NL * H <= 32 && (H == 8 || H == 16 || H == 32)) {
        float *me = seq_sm + wid * pl.total;
        if (H == 8) gru_wave_backward<8, true>(s, w_sm, me, pl, T1, lane);
        else if (H == 16) gru_wave_backward<16, true>(s, w_sm, me, pl, T1, lane);
        else gru_wave_backward<32, false>(s, w_sm, me, pl, T1, lane);
    } else if (wid < n_seq && NL * H <= 32) {
        // wavefront over (layer, step): lane (l, j); the top layer leads, layer l runs step t one stage
        // after layer l + 1 has produced its input gradient for that step.  d h_t^l = recurrent part
        // (dh, from cell (l, t+1)) + input gradient of the layer above (dx, from cell (l+1, t)).
        float *me = seq_sm + wid * pl.total;
        float *dh = me + pl.off_dh, *dhd = me + pl.off_dhd, *dx = me + pl.off_dx;
        const int l = lane / H, j = lane - l * H;
        const bool valid = lane < NL * H;
        int wbase = 0;
        for (int i = 0; i < l && valid; ++i) wbase += 3 * H * gru_row_stride(s, i);
        const int in = valid ? gru_in(s, l) : 0, rs = valid ? gru_row_stride(s, l) : 1;
        const float *wcol = w_sm + wbase;
#pragma unroll 1
        for (int st = 0; st < T1 + NL - 1; ++st) {
            const int t = (T1 - 1) - (st - (NL - 1 - l));
            const bool active = valid && t >= 0 && t <= T1 - 1;
            float *dg = me + pl.off_dg + (t * NL + l) * 4 * H;
            if (active) {
                const float *sv = me + pl.off_sv + (t * NL + l) * 4 * H;
                const float hprev = t > 0 ? me[pl.off_hn + ((t - 1) * NL + l) * H + j] : me[pl.off_h0 + l * H + j];
                const float r = sv[j], z = sv[H + j], n = sv[2 * H + j], ghn = sv[3 * H + j];
                const float d = dh[l * H + j] + dx[l * H + j];
                const float dz = d * (hprev - n), dn = d * (1.f - z);
                const float dan = dn * (1.f - n * n);
                const float dr = dan * ghn;
                dg[j] = dr * (r * (1.f - r));
                dg[H + j] = dz * (z * (1.f - z));
                dg[2 * H + j] = dan;
                dg[3 * H + j] = dan * r;
                dhd[l * H + j] = d * z;
            }
            __syncwarp();
            float sh = 0.f, sx = 0.f;
            if (active) {
                sh = dhd[l * H + j];
#pragma unroll 4
                for (int g = 0; g < 2 * H; ++g) sh = fmaf(wcol[g * rs + in + j], dg[g], sh);
#pragma unroll 4
                for (int g = 2 * H; g < 3 * H; ++g) sh = fmaf(wcol[g * rs + in + j], dg[g + H], sh);
                if (l > 0) {
#pragma unroll 4
                    for (int g = 0; g < 3 * H; ++g) sx = fmaf(wcol[g * rs + j], dg[g], sx);
                }
            }
            __syncwarp();
            if (active) {
                dh[l * H + j] = sh;
                if (l > 0) dx[(l - 1) * H + j] = sx;
            }
            __syncwarp();
        }
    } else if (wid < n_seq) {
        float *me = seq_sm + wid * pl.total;
        float *dh = me + pl.off_dh, *dhd = me + pl.off_dhd;
#pragma unroll 1
        for (int t = T1 - 1; t >= 0; --t) {
#pragma unroll 1
            for (int l = NL - 1; l >= 0; --l) {
                const int in = gru_in(s, l), rs = gru_row_stride(s, l);
                int wbase = 0;
                for (int i = 0; i < l; ++i) wbase += 3 * H * gru_row_stride(s, i);
                const float *sv = me + pl.off_sv + (t * NL + l) * 4 * H;
                float *dg = me + pl.off_dg + (t * NL + l) * 4 * H;
                const float *hprev = t > 0 ? me + pl.off_hn + ((t - 1) * NL + l) * H : me + pl.off_h0 + l * H;
                for (int j = lane; j < H; j += 32) {
                    const float r = sv[j], z = sv[H + j], n = sv[2 * H + j], ghn = sv[3 * H + j];
                    const float d = dh[l * H + j];
                    // h' = (h - n) * z + n
                    const float dz = d * (hprev[j] - n), dn = d * (1.f - z);
                    const float dan = dn * (1.f - n * n);       // tanh
                    const float dr = dan * ghn;
                    dg[j] = dr * (r * (1.f - r));               // d pre-activation of r
                    dg[H + j] = dz * (z * (1.f - z));           // d pre-activation of z
                    dg[2 * H + j] = dan;                        // d (W_in x + b_in)
                    dg[3 * H + j] = dan * r;                    // d (W_hn h + b_hn)
                    dhd[j] = d * z;
                }
                __syncwarp();
                for (int k = lane; k < H; k += 32) {
                    float sh = dhd[k];
                    for (int g = 0; g < 3 * H; ++g) {
                        const float dgh = g < 2 * H ? dg[g] : dg[g + H];
                        sh = fmaf(w_sm[wbase + g * rs + in + k], dgh, sh);
                    }
                    if (l > 0) {  // input of this layer = output of layer l-1 at the same step
                        float sx = 0.f;
                        for (int g = 0; g < 3 * H; ++g) sx = fmaf(w_sm[wbase + g * rs + k], dg[g], sx);
                        dh[(l - 1) * H + k] += sx;
                    }
                    dh[l * H + k] = sh;  // gradient of this layer's output at step t-1
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // ---- partial weight gradients, one row of grad_part per SEQUENCE (summed in sequence order by the
    // reduce + Adam kernel).  A work item is (sequence, layer, W_ih | W_hh, gate row, chunk of 8 columns):
    // the gate gradient of a step is read once for 8 FMAs, and the row's bias gradient rides along.
    {
        constexpr int CH = 8;
        int per_seq = 0;
        for (int l = 0; l < NL; ++l) per_seq += 3 * H * ((gru_in(s, l) + CH - 1) / CH + (H + CH - 1) / CH);
        for (int item = threadIdx.x; item < n_seq * per_seq; item += blockDim.x) {
            const int w = item / per_seq;
            int r = item - w * per_seq, l = 0;
            for (;; ++l) {
                const int n_l = 3 * H * ((gru_in(s, l) + CH - 1) / CH + (H + CH - 1) / CH);
                if (r < n_l) break;
                r -= n_l;
            }
            const int in = gru_in(s, l), c_ih = (in + CH - 1) / CH, c_hh = (H + CH - 1) / CH;
            const bool hh = r >= 3 * H * c_ih;
            if (hh) r -= 3 * H * c_ih;
            const int nch = hh ? c_hh : c_ih, g = r / nch, c0 = (r - g * nch) * CH;
            const int K = hh ? H : in, nc = min(CH, K - c0);
            const int gi = (hh && g >= 2 * H) ? g + H : g;  // W_hh / b_hh see d(W_hn h + b_hn) for the n gate
            const float *me = seq_sm + w * pl.total;
            float acc[CH], accb = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[c] = 0.f;
#pragma unroll 2
            for (int t = 0; t < T1; ++t) {
                const float d = me[pl.off_dg + (t * NL + l) * 4 * H + gi];
                const float *v = !hh ? (l == 0 ? me + pl.off_xs + t * in0 : me + pl.off_hn + (t * NL + l - 1) * H)
                                     : (t > 0 ? me + pl.off_hn + ((t - 1) * NL + l) * H : me + pl.off_h0 + l * H);
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    if (c < nc) acc[c] = fmaf(d, v[c0 + c], acc[c]);
                accb += d;
            }
            float *gl = a.grad_part + (seq0 + w) * a.part_stride + gru_layer_off(s, l);
            float *gw = gl + (hh ? 3 * H * in + g * H : g * in) + c0;
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (c < nc) gw[c] = acc[c];
            if (c0 == 0) gl[3 * H * in + 3 * H * H + (hh ? 3 * H : 0) + g] = accb;
        }
    }
}

}  // namespace asac

using namespace asac;

static const int kSmemLimitRep = 227 * 1024;

static int validate_gru(const AsacGruShape *s) {
    ASAC_REQUIRE(s != nullptr, "null GRU shape");
    ASAC_REQUIRE(s->obs_size > 0 && s->action_size >= 0 && s->hidden > 0 && s->layers > 0, "bad GRU shape");
    ASAC_UNSUPPORTED(s->hidden > 64, "GRU width %d > 64", s->hidden);
    ASAC_UNSUPPORTED(s->layers > ASAC_GRU_MAX_LAYERS, "GRU layers %d > %d", s->layers, ASAC_GRU_MAX_LAYERS);
    ASAC_UNSUPPORTED(s->obs_size + s->action_size > 256, "GRU input width %d > 256", s->obs_size + s->action_size);
    return ASAC_OK;
}

extern "C" int64_t asac_gru_param_count(const AsacGruShape *s) {
    if (validate_gru(s) != ASAC_OK) return -1;
    return gru_count(*s);
}

extern "C" int asac_gru_backward_tile(const AsacGruShape *s, int t_grad) {
    if (validate_gru(s) != ASAC_OK) return ASAC_EINVAL;
    if (t_grad < 0) return ASAC_EINVAL;
    const int per = gru_bwd_plan(*s, t_grad + 1).total, fixed = gru_weight_floats(*s);
    for (int tile = 4; tile >= 1; tile >>= 1)
        if ((fixed + tile * per) * 4 <= kSmemLimitRep) return tile;
    set_error("asac_gru_backward_tile: %d steps of a %d x %d GRU do not fit in shared memory", t_grad + 1, s->layers,
              s->hidden);
    return ASAC_EUNSUPPORTED;
}

template <typename K>
static int grant_smem(K kernel, int bytes, int *granted, const char *name) {
    ASAC_UNSUPPORTED(bytes > kSmemLimitRep, "%s needs %d bytes of shared memory (> %d)", name, bytes, kSmemLimitRep);
    int dev = 0;
    ASAC_CUDA(cudaGetDevice(&dev));
    if (bytes > 48 * 1024 && (dev >= 16 || granted[dev] < bytes)) {
        ASAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev < 16) granted[dev] = bytes;
    }
    return ASAC_OK;
}

extern "C" int asac_gru_forward(const AsacGruShape *s, const AsacGruNet *nets, int n_nets, const float *obs,
                                const float *actions, int bn_stride, const float *pre_actions, const float *h0,
                                int64_t h0_b_stride, int batch, int seq_len, void *stream) {
    int rc = validate_gru(s);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(nets && (n_nets == 1 || n_nets == 2) && obs && batch > 0 && seq_len > 0, "asac_gru_forward: bad arguments");
    ASAC_REQUIRE(s->action_size == 0 || pre_actions || actions || seq_len == 1, "asac_gru_forward: no actions given");
    GruFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.s = *s;
    for (int i = 0; i < n_nets; ++i) {
        ASAC_REQUIRE(nets[i].params && nets[i].states, "asac_gru_forward: net %d lacks params / states", i);
        a.net[i] = nets[i];
    }
    a.obs = obs; a.actions = actions; a.pre_actions = pre_actions; a.h0 = h0; a.h0_b_stride = h0_b_stride;
    a.bn_stride = bn_stride; a.batch = batch; a.seq_len = seq_len;
    const int bytes = (gru_weight_floats(*s) + GRU_FWD_WARPS * gru_fwd_warp_floats(*s, seq_len)) * 4;
    static thread_local int granted[16];
    if ((rc = grant_smem(k_gru_forward, bytes, granted, "k_gru_forward")) != ASAC_OK) return rc;
    ASAC_CUDA(launch_ex(k_gru_forward, dim3((batch + GRU_FWD_WARPS - 1) / GRU_FWD_WARPS, n_nets), dim3(GRU_FWD_WARPS * 32),
                        (size_t)bytes, (cudaStream_t)stream, 0, false, a));
    ASAC_LAUNCHED("k_gru_forward");
    return ASAC_OK;
}

extern "C" int asac_gru_backward(const AsacGruShape *s, const float *params, const float *obs, const float *actions,
                                 int bn_stride, const float *pre_actions, const float *h0, int64_t h0_b_stride,
                                 int batch, int seq_len, int t_grad, const float *grad_state, int ensemble,
                                 const float *hn, const float *save, float *grad_part, void *stream) {
    int rc = validate_gru(s);
    if (rc != ASAC_OK) return rc;
    ASAC_REQUIRE(params && obs && grad_state && hn && save && grad_part, "asac_gru_backward: null pointer");
    ASAC_REQUIRE(batch > 0 && t_grad >= 0 && t_grad < seq_len && ensemble >= 1, "asac_gru_backward: bad sizes");
    const int tile = asac_gru_backward_tile(s, t_grad);
    if (tile < 1) return tile;
    GruBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.s = *s;
    a.params = params; a.obs = obs; a.actions = actions; a.pre_actions = pre_actions; a.h0 = h0;
    a.grad_state = grad_state; a.hn = hn; a.save = save; a.grad_part = grad_part;
    a.h0_b_stride = h0_b_stride; a.part_stride = (gru_count(*s) + 3) / 4 * 4;
    a.bn_stride = bn_stride; a.batch = batch; a.seq_len = seq_len; a.t_grad = t_grad; a.ensemble = ensemble;
    a.tile = tile;
    const int bytes = (gru_weight_floats(*s) + tile * gru_bwd_plan(*s, t_grad + 1).total) * 4;
    static thread_local int granted[16];
    if ((rc = grant_smem(k_gru_backward, bytes, granted, "k_gru_backward")) != ASAC_OK) return rc;
    ASAC_CUDA(launch_ex(k_gru_backward, dim3((batch + tile - 1) / tile), dim3(GRU_BWD_THREADS), (size_t)bytes,
                        (cudaStream_t)stream, 0, false, a));
    ASAC_LAUNCHED("k_gru_backward");
    return ASAC_OK;
}
