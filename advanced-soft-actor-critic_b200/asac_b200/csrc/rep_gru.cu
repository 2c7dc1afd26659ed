// Recurrent representation kernels (sm_100a): the stock multi-layer GRU wrapper of the reference
// (algorithm/nn_models/layers/seq_layers.py:14-114, no padding mask) as used by a plugin ModelRep of
// the form of envs/test/nn_rnn.py:6-21 — forward over whole sampled windows and back-propagation
// through time of the critic loss's state gradient (sac_base.py:1510-1601, 2066-2105).
//
// The recurrence is a chain of L x layers dependent cells of a few hundred FMAs each: there is no
// parallelism inside a sequence worth a block barrier, so ONE WARP owns one sequence (lanes = gate
// rows in the matrix-vector phase, = hidden units in the gate phase, __syncwarp between them) and
// the batch spreads over the SMs.  Weights sit in shared memory in rows of odd stride
// [W_ih row | W_hh row | b_ih | b_hh] (conflict-free for row-per-lane and column-per-lane reads);
// the sequence's inputs are staged once, coalesced, before the chain starts.
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

// flat parameters of every layer -> padded shared-memory rows (whole CTA)
__device__ void gru_stage_weights(float *w_sm, const float *params, const AsacGruShape &s) {
    const int H = s.hidden;
    int base = 0;
    for (int l = 0; l < s.layers; ++l) {
        const int in = gru_in(s, l), rs = gru_row_stride(s, l);
        const float *p = params + gru_layer_off(s, l);
        const int n_ih = 3 * H * in, n_hh = 3 * H * H;
        for (int i = threadIdx.x; i < n_ih + n_hh + 6 * H; i += blockDim.x) {
            int g, c;
            if (i < n_ih) { g = i / in; c = i - g * in; }
            else if (i < n_ih + n_hh) { const int j = i - n_ih; g = j / H; c = in + (j - g * H); }
            else if (i < n_ih + n_hh + 3 * H) { g = i - n_ih - n_hh; c = in + H; }
            else { g = i - n_ih - n_hh - 3 * H; c = in + H + 1; }
            w_sm[base + g * rs + c] = __ldg(p + i);
        }
        base += 3 * H * rs;
    }
}

// x_t = [obs[b, t], pre_action[b, t]] for t < T into xs[T][in0] (one warp)
__device__ __forceinline__ void gru_stage_inputs(float *xs, const float *obs, const float *actions, int bn_stride,
                                                 const float *pre_actions, int64_t seq, int L, int T, int So, int A,
                                                 int lane) {
    const int in0 = So + A;
    for (int i = lane; i < T * in0; i += 32) {
        const int t = i / in0, c = i - t * in0;
        float v;
        if (c < So) v = obs[(seq * L + t) * So + c];
        else if (pre_actions) v = pre_actions[(seq * L + t) * A + (c - So)];
        else v = t > 0 ? actions[(seq * bn_stride + t - 1) * A + (c - So)] : 0.f;  // operators.py:39-59
        xs[i] = v;
    }
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

__host__ __device__ __forceinline__ int gru_fwd_warp_floats(const AsacGruShape &s, int L) {
    return ((L * (s.obs_size + s.action_size) + s.layers * s.hidden + 4 * s.hidden) + 3) & ~3;
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
    float *hbuf = xs + L * in0, *gates = hbuf + NL * H;
    if (seq < a.batch) {
        gru_stage_inputs(xs, a.obs, a.actions, a.bn_stride, a.pre_actions, seq, L, L, s.obs_size, s.action_size, lane);
        for (int i = lane; i < NL * H; i += 32) hbuf[i] = a.h0 ? a.h0[seq * a.h0_b_stride + i] : 0.f;
    }
    __syncthreads();
    if (seq >= a.batch) return;
#pragma unroll 1
    for (int t = 0; t < L; ++t) {
        int wbase = 0;
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            const int in = gru_in(s, l), rs = gru_row_stride(s, l);
            const float *x = l == 0 ? xs + t * in0 : hbuf + (l - 1) * H;  // layer l-1's output of this step
            float *hp = hbuf + l * H;
            for (int g = lane; g < 3 * H; g += 32) {
                const float *wr = w_sm + wbase + g * rs;
                float gi = wr[in + H], gh = wr[in + H + 1];
                for (int k = 0; k < in; ++k) gi = fmaf(wr[k], x[k], gi);
                for (int k = 0; k < H; ++k) gh = fmaf(wr[in + k], hp[k], gh);
                if (g < 2 * H) {
                    gates[g] = gh + gi;
                } else {
                    gates[g] = gi;
                    gates[g + H] = gh;
                }
            }
            __syncwarp();
            for (int j = lane; j < H; j += 32) {
                const float r = sigmoidf_(gates[j]), z = sigmoidf_(gates[H + j]);
                const float ghn = gates[3 * H + j];
                const float n = tanhf(gates[2 * H + j] + ghn * r);
                const float hnew = (hp[j] - n) * z + n;
                hp[j] = hnew;
                const int64_t cell = (seq * L + t) * NL + l;
                if (net.hn) net.hn[cell * H + j] = hnew;
                if (net.save) {
                    float *sv = net.save + cell * 4 * H;
                    sv[j] = r; sv[H + j] = z; sv[2 * H + j] = n; sv[3 * H + j] = ghn;
                }
                if (l == NL - 1) net.states[(seq * L + t) * H + j] = hnew;
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
    int off_xs, off_hn, off_h0, off_sv, off_dg, off_dh, off_dhd, total;
};
__host__ __device__ __forceinline__ GruBwdPlan gru_bwd_plan(const AsacGruShape &s, int T1) {
    GruBwdPlan p;
    const int H = s.hidden, NL = s.layers;
    int o = 0;
    p.off_xs = o; o += T1 * (s.obs_size + s.action_size);
    p.off_hn = o; o += T1 * NL * H;
    p.off_h0 = o; o += NL * H;
    p.off_sv = o; o += T1 * NL * 4 * H;
    p.off_dg = o; o += T1 * NL * 4 * H;
    p.off_dh = o; o += NL * H;
    p.off_dhd = o; o += H;
    p.total = (o + 3) & ~3;
    return p;
}

// grid ceil(B / tile); warp w < tile runs the BPTT of sequence blockIdx.x * tile + w, then the whole CTA
// turns the stored gate gradients into this tile's partial weight gradients.
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
    if (wid < n_seq) {
        const int64_t seq = seq0 + wid;
        float *me = seq_sm + wid * pl.total;
        gru_stage_inputs(me + pl.off_xs, a.obs, a.actions, a.bn_stride, a.pre_actions, seq, L, T1, s.obs_size,
                         s.action_size, lane);
        for (int i = lane; i < T1 * NL * H; i += 32) me[pl.off_hn + i] = a.hn[seq * L * NL * H + i];
        for (int i = lane; i < NL * H; i += 32) me[pl.off_h0 + i] = a.h0 ? a.h0[seq * a.h0_b_stride + i] : 0.f;
        for (int i = lane; i < T1 * NL * 4 * H; i += 32) me[pl.off_sv + i] = a.save[seq * L * NL * 4 * H + i];
        for (int i = lane; i < NL * H; i += 32) {
            float g = 0.f;
            if (i >= (NL - 1) * H)  // d loss / d state[:, t_grad], summed over the critics in member order
                for (int e = 0; e < a.ensemble; ++e) g += a.grad_state[((int64_t)e * a.batch + seq) * H + (i - (NL - 1) * H)];
            me[pl.off_dh + i] = g;
        }
    }
    __syncthreads();
    if (wid < n_seq) {
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
    // ---- partial weight gradients of this tile: sum over its sequences and steps, in that order
    float *gout = a.grad_part + (int64_t)blockIdx.x * a.part_stride;
    for (int l = 0; l < NL; ++l) {
        const int in = gru_in(s, l);
        const int n_ih = 3 * H * in, n_hh = 3 * H * H;
        float *gl = gout + gru_layer_off(s, l);
        for (int i = threadIdx.x; i < n_ih + n_hh + 6 * H; i += blockDim.x) {
            int g, c, kind;  // kind 0: W_ih, 1: W_hh, 2: b_ih, 3: b_hh
            if (i < n_ih) { g = i / in; c = i - g * in; kind = 0; }
            else if (i < n_ih + n_hh) { const int j = i - n_ih; g = j / H; c = j - g * H; kind = 1; }
            else if (i < n_ih + n_hh + 3 * H) { g = i - n_ih - n_hh; c = 0; kind = 2; }
            else { g = i - n_ih - n_hh - 3 * H; c = 0; kind = 3; }
            const bool hh = (kind & 1) != 0;
            const int gi = (hh && g >= 2 * H) ? g + H : g;  // W_hh / b_hh see d(W_hn h + b_hn) for the n gate
            float acc = 0.f;
            for (int w = 0; w < n_seq; ++w) {
                const float *me = seq_sm + w * pl.total;
                for (int t = 0; t < T1; ++t) {
                    const float d = me[pl.off_dg + (t * NL + l) * 4 * H + gi];
                    float v = 1.f;
                    if (kind == 0) v = l == 0 ? me[pl.off_xs + t * in0 + c] : me[pl.off_hn + (t * NL + l - 1) * H + c];
                    else if (kind == 1) v = t > 0 ? me[pl.off_hn + ((t - 1) * NL + l) * H + c] : me[pl.off_h0 + l * H + c];
                    acc = fmaf(d, v, acc);
                }
            }
            gl[i] = acc;
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
