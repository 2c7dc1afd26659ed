"""ctypes binding of ``libasac_b200.so`` (the C ABI declared in ``include/asac_b200.h``).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There is
no CPU fallback: a missing library raises at import of any product module that needs it,
and a missing CUDA device raises on the first call.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
# ASAC_B200_LIB: alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = Path(os.environ['ASAC_B200_LIB']) if os.environ.get('ASAC_B200_LIB') else PKG_DIR / 'libasac_b200.so'
CSRC_DIR = PKG_DIR / 'csrc'

MAX_COLUMNS = 16
MAX_BRANCHES = 8
MAX_NSTEP = 16
MAX_DEPTH = 4
MAX_ENSEMBLE = 8

ROLE_COPY, ROLE_INDEX, ROLE_ACTION, ROLE_REWARD, ROLE_DONE, ROLE_MU_PROB, ROLE_HIDDEN = range(7)

vp = C.c_void_p


class AsacColumn(C.Structure):
    _fields_ = [('ring', vp), ('out', vp), ('row_bytes', C.c_int32), ('out_stride', C.c_int32),
                ('out_offset', C.c_int32), ('role', C.c_int32)]


class AsacColumnTable(C.Structure):
    _fields_ = [('n_columns', C.c_int32), ('index_column', C.c_int32), ('col', AsacColumn * MAX_COLUMNS)]


class AsacSacConfig(C.Structure):
    _fields_ = [('learning_rate', C.c_double),
                ('batch', C.c_int32), ('seq_len', C.c_int32), ('burn_in', C.c_int32), ('n_step', C.c_int32),
                ('state_size', C.c_int32), ('action_size', C.c_int32), ('ensemble', C.c_int32),
                ('q_hidden', C.c_int32), ('q_depth', C.c_int32), ('pi_hidden', C.c_int32), ('pi_depth', C.c_int32),
                ('use_n_step_is', C.c_int32), ('use_priority', C.c_int32), ('use_auto_alpha', C.c_int32),
                ('update_target_per_step', C.c_int32), ('bn_stride', C.c_int32),
                ('tau', C.c_float), ('one_minus_tau', C.c_float), ('gamma', C.c_float), ('v_rho', C.c_float),
                ('v_c', C.c_float), ('clip_epsilon', C.c_float), ('target_c_alpha', C.c_float),
                ('td_error_min', C.c_float), ('td_error_max', C.c_float), ('per_alpha', C.c_float),
                ('gamma_ratio', C.c_float * MAX_NSTEP), ('lambda_ratio', C.c_float * MAX_NSTEP),
                ('rep_kind', C.c_int32), ('rep_param_stride', C.c_int32), ('ensemble_sample', C.c_int32)]


class AsacSacParams(C.Structure):
    _fields_ = [('q', vp), ('q_target', vp), ('pi', vp), ('log_alpha', vp),
                ('q_m', vp), ('q_v', vp), ('pi_m', vp), ('pi_v', vp), ('alpha_m', vp), ('alpha_v', vp),
                ('counters', vp)]


class AsacSacBatch(C.Structure):
    _fields_ = [('states', vp), ('actions', vp), ('rewards', vp), ('dones', vp), ('last_masks', vp),
                ('padding_masks', vp), ('mu_probs', vp), ('priority_is', vp),
                ('eps_y', vp), ('eps_pi', vp), ('eps_alpha', vp), ('eps_td', vp),
                ('states_post', vp), ('target_states', vp), ('ensemble_perms', vp)]


class AsacWriteColumn(C.Structure):
    _fields_ = [('ring', vp), ('rows', vp), ('row_bytes', C.c_int64)]


class AsacWriteTable(C.Structure):
    _fields_ = [('n_columns', C.c_int32), ('col', AsacWriteColumn * MAX_COLUMNS)]


MAX_PEERS = 16


class AsacPeerTable(C.Structure):
    _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('recv', vp * MAX_PEERS), ('recv_words', C.c_int64)]


class AsacSacWork(C.Structure):
    _fields_ = [('n_tiles', C.c_int32),
                ('y', vp), ('tq', vp), ('q_val', vp), ('loss_q', vp), ('grad_q_part', vp), ('grad_q', vp),
                ('grad_pi_part', vp), ('grad_pi', vp), ('stats_pi', vp), ('grad_alpha_part', vp),
                ('grad_alpha', vp), ('pi_probs', vp), ('post_parts', vp), ('y_td', vp), ('td_error', vp),
                ('grad_state', vp)]


GRU_MAX_LAYERS = 4


class AsacDiscreteConfig(C.Structure):
    _fields_ = [('branches', C.c_int32), ('sizes', C.c_int32 * MAX_BRANCHES), ('hidden', C.c_int32),
                ('depth', C.c_int32), ('state_size', C.c_int32), ('target_d_alpha', C.c_float),
                ('entropy_penalty', C.c_float)]


class AsacGruShape(C.Structure):
    _fields_ = [('obs_size', C.c_int32), ('action_size', C.c_int32), ('hidden', C.c_int32), ('layers', C.c_int32)]


class AsacGruNet(C.Structure):
    _fields_ = [('params', vp), ('states', vp), ('hn', vp), ('save', vp)]


class AsacGruRep(C.Structure):
    _fields_ = [('shape', AsacGruShape), ('params', vp), ('params_target', vp), ('m', vp), ('v', vp),
                ('obs', vp), ('h0', vp), ('h0_b_stride', C.c_int64),
                ('states', vp), ('states_post', vp), ('target_states', vp), ('hn', vp), ('hn_post', vp),
                ('save', vp), ('grad_part', vp), ('grad', vp), ('rep_tiles', C.c_int32), ('reserved_', C.c_int32)]


i32, i64, u64, f32 = C.c_int, C.c_int64, C.c_uint64, C.c_float
P = C.POINTER

# name -> (restype, argtypes); every symbol include/asac_b200.h declares
PROTOTYPES = {
    'asac_last_error': (C.c_char_p, []),
    'asac_version': (i32, []),
    'asac_launch_count': (i64, []),
    'asac_set_pdl': (i32, [i32]),
    'asac_reset_launch_count': (None, []),
    'asac_tree_update': (i32, [vp, i64, vp, vp, i64, vp]),
    'asac_tree_rebuild': (i32, [vp, i64, vp]),
    'asac_tree_leaf_max': (i32, [vp, i64, vp, vp]),
    'asac_tree_sample': (i32, [vp, i64, i32, vp, u64, vp, vp, vp, vp]),
    'asac_per_sample': (i32, [vp, i64, vp, i32, vp, u64, vp, vp, vp, vp, vp, vp, vp]),
    'asac_per_shard_weights': (i32, [vp, i32, vp, vp, vp, vp, vp]),
    'asac_per_update': (i32, [vp, i64, vp, vp, vp, i32, f32, f32, f32, i32, vp, vp]),
    'asac_per_add': (i32, [vp, i64, vp, i64, i64, vp, i32, vp]),
    'asac_storage_write_rows': (i32, [vp, i64, i64, vp, i64, i64, vp]),
    'asac_storage_write_table': (i32, [P(AsacWriteTable), i64, i64, i64, vp]),
    'asac_ingest_create': (i32, [P(vp), i64, i32, P(vp), P(i64), vp, vp, vp, vp]),
    'asac_ingest_add': (i32, [vp, P(vp), i64, i64, i32, i32, vp]),
    'asac_ingest_row_bytes': (i64, [vp]),
    'asac_ingest_destroy': (None, [vp]),
    'asac_storage_gather': (i32, [P(AsacColumnTable), i64, vp, i32, i32, i32, vp, vp, vp]),
    'asac_storage_scatter': (i32, [vp, i64, vp, vp, i32, i32, i32, vp, i64, i64, vp, i64, vp]),
    'asac_sac_tile_batch': (i32, [P(AsacSacConfig)]),
    'asac_mlp_param_count': (i64, [i32, i32, i32, i32]),
    'asac_mlp_param_stride': (i64, [i32, i32, i32, i32]),
    'asac_dnets_member_floats': (i64, [P(AsacDiscreteConfig)]),
    'asac_dnets_tiles': (i32, [i32]),
    'asac_dnets_forward': (i32, [P(AsacDiscreteConfig), vp, i64, i32, vp, i64, i32, vp, vp]),
    'asac_dnets_backward': (i32, [P(AsacDiscreteConfig), vp, i64, i32, vp, i64, i32, vp, vp, vp, vp]),
    'asac_d_target': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'asac_d_target_dqn': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'asac_d_q_grad': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp, f32, vp, vp, vp, vp]),
    'asac_d_pi_grad': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    'asac_d_probs': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp]),
    'asac_d_alpha': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp, vp, f32, vp, vp, vp]),
    'asac_d_td': (i32, [P(AsacSacConfig), P(AsacDiscreteConfig), vp, vp, vp, vp, i32, vp]),
    'asac_bump_counters': (i32, [vp, i32, vp]),
    'asac_ensemble_perms': (i32, [vp, i32, i32, u64, vp, vp]),
    'asac_sac_value_pass_on_tc': (i32, [P(AsacSacConfig), i32]),
    'asac_sac_polyak': (i32, [P(AsacSacConfig), P(AsacSacParams), f32, vp]),
    'asac_sac_target_y': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), vp]),
    'asac_sac_q_backward': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), vp]),
    'asac_sac_policy_backward': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), vp]),
    'asac_sac_post': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), vp]),
    'asac_sac_reduce_grads': (i32, [P(AsacSacConfig), P(AsacSacWork), i32, vp]),
    'asac_sac_adam': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacWork), i32, f32, vp]),
    'asac_sac_reduce_adam': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacWork), i32, vp]),
    'asac_sac_td_error': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacWork), vp]),
    'asac_sac_advance_step': (i32, [P(AsacSacParams), vp]),
    'asac_sac_step': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), vp]),
    'asac_sac_step_networks': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork), i32,
                                     P(AsacPeerTable), vp]),
    'asac_sac_finish_step': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacWork), vp, i64, vp, vp, vp,
                                   P(AsacPeerTable), vp]),
    'asac_peer_recv_words': (i64, [P(AsacSacConfig), i32]),
    'asac_peer_timeouts': (i32, [i32]),
    'asac_gru_param_count': (i64, [P(AsacGruShape)]),
    'asac_gru_backward_tile': (i32, [P(AsacGruShape), i32]),
    'asac_gru_forward': (i32, [P(AsacGruShape), P(AsacGruNet), i32, vp, vp, i32, vp, vp, i64, i32, i32, vp]),
    'asac_gru_backward': (i32, [P(AsacGruShape), vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, vp, i32, vp, vp, vp,
                                vp]),
    'asac_flat_reduce_adam': (i32, [vp, vp, vp, vp, i32, i64, i64, vp, vp, C.c_double, vp]),
    'asac_flat_polyak': (i32, [vp, vp, i64, vp, i32, f32, f32, i32, vp]),
    'asac_sac_step_networks_rep': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacBatch), P(AsacSacWork),
                                         P(AsacGruRep), i32, P(AsacPeerTable), vp]),
    'asac_sac_staged_tail': (i32, [P(AsacSacConfig), P(AsacSacParams), P(AsacSacWork), vp]),
    'asac_fill_normal': (i32, [vp, i64, u64, vp, i32, vp]),
    'asac_l2_prefetch': (i32, [P(vp), P(i64), i32, vp]),
    'asac_mlp_forward': (i32, [vp, i32, i32, i32, i32, vp, i64, vp, vp]),
    'asac_mlp_forward_tc': (i32, [vp, i32, i32, i32, i32, vp, i64, vp, vp]),
    'asac_mlp_forward_tcf': (i32, [vp, i32, i32, i32, i32, vp, i64, vp, i32, i32, vp]),
    'asac_mlp_forward_tcf_probe': (i32, [vp, i32, i32, i32, i32, vp, i64, vp, i32, i32, vp, vp]),
    'asac_debug_phase_clocks': (i32, [vp]),
    'asac_debug_global_stamps': (i32, [vp]),
    'asac_policy_act': (i32, [vp, i32, i32, i32, i32, vp, i64, vp, vp, i32, u64, vp, vp, vp, vp, i32, vp]),
}


class AsacError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Loads the shared library and binds every prototype.  Raises (no fallback) when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AsacError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                        f'or `make -C {CSRC_DIR}` (there is no CPU fallback)')
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError when the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = load().asac_last_error().decode(errors='replace')
        kind = {-1: 'invalid argument', -2: 'unsupported configuration'}.get(rc, f'CUDA error {rc}')
        if rc == -2:
            raise NotImplementedError(f'{what}: {kind}: {msg}')
        raise AsacError(f'{what}: {kind}: {msg}')


def ptr(t) -> int | None:
    """data_ptr of a CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AsacError('libasac_b200 takes CUDA tensors only (no CPU fallback)')
    if not t.is_contiguous():
        raise AsacError('libasac_b200 takes contiguous tensors')
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
