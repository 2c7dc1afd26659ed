"""Shared helpers for the parity tests (golden loading, oracle construction)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / 'golden'


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}


def sac_case_meta(g: dict) -> dict:
    S, A, E, hidden, depth, B, b, n, steps, use_priority = [int(x) for x in g['meta'][:10]]
    m = dict(S=S, A=A, E=E, hidden=hidden, depth=depth, B=B, b=b, n=n, steps=steps,
             use_priority=bool(use_priority))
    if len(g['meta']) > 10:  # recurrent fixtures: S above is the observation width, the state is the GRU width
        m.update(So=S, rep_layers=int(g['meta'][10]), S=int(g['meta'][11]))
    return m


def sac_hyper_from_golden(g: dict):
    from oracle.sac_oracle import SacHyper
    m = sac_case_meta(g)
    hp = {k[3:]: float(v) for k, v in g.items() if k.startswith('hp.') and np.ndim(v) == 0}
    return SacHyper(state_size=m['S'], action_size=m['A'], ensemble_q_num=m['E'], hidden=m['hidden'],
                    q_depth=m['depth'], policy_depth=m['depth'], burn_in_step=m['b'], n_step=m['n'],
                    tau=hp['tau'], update_target_per_step=int(hp['update_target_per_step']),
                    init_log_alpha=hp['init_log_alpha'], use_auto_alpha=bool(hp['use_auto_alpha']),
                    target_c_alpha=hp['target_c_alpha'], learning_rate=hp['learning_rate'],
                    gamma=hp['gamma'], v_lambda=hp['v_lambda'], v_rho=hp['v_rho'], v_c=hp['v_c'],
                    clip_epsilon=hp['clip_epsilon'], use_n_step_is=bool(hp['use_n_step_is']),
                    use_priority=m['use_priority'],
                    ensemble_q_sample=int(g['ensemble_q_sample']) if 'ensemble_q_sample' in g else 0)


def golden_params(g: dict, prefix: str, E: int):
    """-> (q list, q_target list, policy dict, log_c_alpha) keyed by torch state_dict names."""
    def sub(tag):
        pre = f'{prefix}.{tag}.'
        return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}
    return ([sub(f'q{i}') for i in range(E)], [sub(f'qt{i}') for i in range(E)], sub('pi'),
            g[f'{prefix}.log_c_alpha'])


def golden_batch(g: dict, step: int):
    from oracle.sac_oracle import SacBatch, SacNoise
    pre = f's{step}.in.'
    t = lambda k: torch.from_numpy(g[pre + k])
    batch = SacBatch(states=t('states'), actions=t('actions'), rewards=t('rewards'), dones=t('dones'),
                     mu_probs=t('mu_probs'), last_masks=t('last_masks'), padding_masks=t('padding_masks'),
                     priority_is=t('priority_is') if (pre + 'priority_is') in g else None)
    noise = SacNoise(eps_y=t('eps_y'), eps_pi=t('eps_pi'), eps_alpha=t('eps_alpha'), eps_td=t('eps_td'))
    return batch, noise


def golden_rep_params(g: dict, prefix: str):
    """-> (rep dict, target rep dict) keyed by the plugin's state_dict names (rnn._grus.<l>.*)."""
    def sub(tag):
        pre = f'{prefix}.{tag}.'
        return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}
    return sub('rep'), sub('rept')


def golden_rep_batch(g: dict, step: int):
    from oracle.rep_oracle import SacRepBatch
    from oracle.sac_oracle import SacNoise
    pre = f's{step}.in.'
    t = lambda k: torch.from_numpy(g[pre + k])
    batch = SacRepBatch(obs=t('obs'), hidden0=t('hidden0'), actions=t('actions'), rewards=t('rewards'),
                        dones=t('dones'), mu_probs=t('mu_probs'), last_masks=t('last_masks'),
                        padding_masks=t('padding_masks'),
                        priority_is=t('priority_is') if (pre + 'priority_is') in g else None)
    noise = SacNoise(eps_y=t('eps_y'), eps_pi=t('eps_pi'), eps_alpha=t('eps_alpha'), eps_td=t('eps_td'))
    return batch, noise


def rep_oracle_from_golden(g: dict, dtype=torch.float32):
    from oracle.rep_oracle import SacRepOracle
    m = sac_case_meta(g)
    oracle = SacRepOracle(sac_hyper_from_golden(g), m['So'], m['rep_layers'], dtype=dtype)
    oracle.load_params(*golden_params(g, 'init', m['E']))
    oracle.load_rep(*golden_rep_params(g, 'init'))
    return oracle


def rel_err(a, b) -> float:
    """max |a-b| / max(1, max|b|): the "1e-5 relative-to-scale" of SURVEY.md §7."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))
