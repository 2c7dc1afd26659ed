"""Representations the fused kernels do not cover, run as the plugin's own torch modules on the device.

The learner's native path knows two representations: ``ModelSimpleRep`` (the gather writes the state) and
the stock ``GRU`` over ``cat[obs, pre_action]`` (``csrc/rep_gru.cu``).  Every other ``ModelRep`` a plugin
file defines — convolutional / ray encoders in front of a GRU, dense layers around it, the episode
attention stack (``seq_encoder=ATTN``), several observations — is arbitrary torch code; for those the
learner keeps EVERYTHING ELSE on its kernels (prioritized replay, window gather, V-trace targets, critic /
policy / alpha updates, Adam, priority update) and calls the plugin's module for the three
``get_l_states`` passes of a step (sac_base.py:2066-2105) with torch autograd for its backward:

    states        = model_rep(...)            (autograd graph kept)
    target_states = model_target_rep(...)     (no grad)
    ... k_value_pass, k_q_backward -> d loss_i / d state[:, burn_in]  (wrk.grad_state) ...
    states.backward(sum_i grad_state_i at column burn_in)             (the representation's share of loss.backward())
    Adam on the representation's flat parameter buffer                (asac_flat_reduce_adam, the learner's kernel)
    states_post, next_hidden = model_rep(...) (no grad, new weights)

The module's parameters (and their ``.grad``) are views of flat fp32 buffers, so the Adam / Polyak kernels,
the checkpoint (``model_rep`` / ``model_target_rep`` / ``optimizer_rep`` keys, sac_base.py:506-509) and the
gradient all-reduce of the data-parallel learner see one contiguous vector.  cuDNN / cuBLAS run the
encoder's own math here: that part of the step is library code, not this repo's kernels — it is the
drop-in's compatibility path, the native one being the two forms above.
"""
from __future__ import annotations

import torch
from torch import nn

from . import lowering


class TorchRepBridge:
    def __init__(self, model_rep: nn.Module, model_target_rep: nn.Module, attn: bool, burn_in_step: int,
                 device: torch.device):
        self.rep, self.target = model_rep, model_target_rep
        self.attn, self.b, self.device = attn, burn_in_step, device
        self.params = [p for p in model_rep.parameters()]
        self.tparams = [p for p in model_target_rep.parameters()]
        if [tuple(p.shape) for p in self.params] != [tuple(p.shape) for p in self.tparams]:
            raise lowering.NotStockNetwork('target representation differs from the online one')
        if any(p.dtype != torch.float32 for p in self.params):
            raise NotImplementedError('representation parameters must be float32')
        self.count = sum(p.numel() for p in self.params)
        self.stride = (self.count + 3) // 4 * 4
        f32 = dict(dtype=torch.float32, device=device)
        self.flat, self.flat_target = torch.zeros(self.stride, **f32), torch.zeros(self.stride, **f32)
        self.m, self.v = torch.zeros(self.stride, **f32), torch.zeros(self.stride, **f32)
        self.grad, self.grad_out = torch.zeros(self.stride, **f32), torch.zeros(self.stride, **f32)
        lowering.bind_parameters(self.params, self.flat)
        lowering.bind_parameters(self.tparams, self.flat_target)
        off = 0
        for p in self.params:  # autograd accumulates into these views: the flat gradient needs no gather
            p.grad = self.grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        # (nn.GRU.flatten_parameters() must NOT run after this: it would move the weights into cuDNN's own buffer)

    # ------------------------------------------------------------------ shapes (sac_base.py:352-386)
    @torch.no_grad()
    def probe(self, obs_shapes, action_size: int, batch: int = 2):
        dev = self.device
        obs = [torch.rand(batch, 1, *shape, device=dev) for shape in obs_shapes]
        pre_action = torch.rand(batch, 1, action_size, device=dev)
        with self._math():
            if self.attn:
                index = torch.zeros((batch, 1), dtype=torch.int32, device=dev)
                state, hidden, _ = self.rep(1, index, obs, pre_action, None)
            else:
                state, hidden = self.rep(obs, pre_action, None)
        return int(state.shape[-1]), tuple(hidden.shape[2:])

    @staticmethod
    def _math():
        """fp32 everywhere: cuDNN convolutions / RNNs default to TF32, the parity bound is 1e-5."""
        return torch.backends.cudnn.flags(enabled=True, allow_tf32=False)

    # ------------------------------------------------------------------ get_l_states (sac_base.py:1118-1157)
    def l_states(self, index, padding_mask, obs_list, pre_action, hidden, target: bool):
        model = self.target if target else self.rep
        with self._math():
            if self.attn:
                state, hidden_out, _ = model(index.shape[1], index, obs_list, pre_action, hidden[:, :1],
                                             is_prev_hidden_state=True, padding_mask=padding_mask)
            else:
                state, hidden_out = model(obs_list, pre_action, hidden, padding_mask=padding_mask)
        return state, hidden_out

    def zero_grad(self) -> None:
        self.grad.zero_()
        off = 0
        for p in self.params:  # a plugin (or torch) may have replaced .grad; the flat views are the contract
            want = self.grad[off:off + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                p.grad = want
            off += p.numel()

    def backward(self, states: torch.Tensor, grad_state_b: torch.Tensor) -> None:
        """d(sum_i mean loss_i) / d rep params: the critics only see state[:, burn_in] (sac_base.py:1516-1601),
        ``grad_state_b`` [B, S] = the sum over the members (and over the discrete heads) of what the critic backward
        kernels wrote."""
        g = torch.zeros_like(states)
        g[:, self.b] = grad_state_b
        self.zero_grad()
        with self._math():
            torch.autograd.backward([states], [g], inputs=self.params)
