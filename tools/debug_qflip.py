"""Debug: per-sample clipped-loss branch of the critic loss, CUDA vs oracle (B=1024, n=5 case)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import numpy as np, torch
from oracle.sac_oracle import SacBatch, SacHyper, SacNoise, SacOracle, q_forward
from tests.cuda_harness import SacCuda
S, A, B, n, depth, seed = 3, 1, 1024, 5, 2, 1
hp = SacHyper(state_size=S, action_size=A, ensemble_q_num=2, hidden=64, q_depth=depth, policy_depth=depth, n_step=n)
oracle = SacOracle(hp, seed=seed)
gen = torch.Generator().manual_seed(seed)
with torch.no_grad():
    for net in oracle.q + oracle.q_target + [oracle.policy]:
        for t in net.values():
            t.add_(torch.randn(t.shape, generator=gen) * 0.03)
L = n + 1
r = lambda *s: torch.randn(*s, generator=gen)
batch = SacBatch(states=r(B, L, S), actions=torch.rand(B, L - 1, A, generator=gen) * 1.8 - 0.9,
                 rewards=r(B, L - 1), dones=torch.rand(B, L - 1, generator=gen) < 0.1,
                 mu_probs=torch.rand(B, L - 1, A, generator=gen) + 0.05,
                 last_masks=torch.rand(B, L - 1, generator=gen) < 0.05,
                 padding_masks=torch.zeros(B, L - 1, dtype=torch.bool),
                 priority_is=torch.rand(B, 1, generator=gen) * 0.9 + 0.1)
noise = SacNoise(eps_y=r(B, n + 1, A), eps_pi=r(B, A), eps_alpha=r(B, A), eps_td=r(B, n + 1, A))
cuda = SacCuda(hp, B)
cb = cuda.make_batch(batch, noise)
cuda.sync_from_oracle(oracle)
oracle.polyak(hp.tau); cuda.polyak(); cuda.sync_from_oracle(oracle, what=('qt',))
with torch.no_grad():
    tq_o = [q_forward(oracle.q_target[i], depth, batch.states[:, 0], batch.actions[:, 0]).numpy().reshape(-1) for i in range(2)]
res = oracle.train_q(batch, noise.eps_y)
cuda.target_y(cb); cuda.q_backward(cb); cuda.reduce_grads(0)
y_o = res['y'].numpy().reshape(-1); y_c = cuda.wk['y'].cpu().numpy()
eps = np.float32(hp.clip_epsilon)
for i in range(2):
    q_o = res['q'][i].numpy().reshape(-1); q_c = cuda.wk['q_val'][i].cpu().numpy(); tq_c = cuda.wk['tq'][i].cpu().numpy()
    def branch(q, tq, y):
        diff = q - tq; cl = np.clip(diff, -eps, eps); cq = tq + cl
        la = (cq - y) ** 2; lb = (q - y) ** 2
        inside = (diff >= -eps) & (diff <= eps)
        return la, lb, inside, np.sign(la - lb)
    la_o, lb_o, in_o, s_o = branch(q_o, tq_o[i], y_o)
    la_c, lb_c, in_c, s_c = branch(q_c, tq_c, y_c)
    d = np.where((s_o != s_c) | (in_o != in_c))[0]
    print(f'net {i}: max|q diff| {np.abs(q_o-q_c).max():.2e} max|tq diff| {np.abs(tq_o[i]-tq_c).max():.2e} max|y diff| {np.abs(y_o-y_c).max():.2e}; branch differs on {len(d)} samples')
    for j in d[:10]:
        print(f'   j={j} q {q_o[j]:.7f}/{q_c[j]:.7f} tq {tq_o[i][j]:.7f}/{tq_c[j]:.7f} y {y_o[j]:.7f}/{y_c[j]:.7f} la-lb {la_o[j]-lb_o[j]:.3e}/{la_c[j]-lb_c[j]:.3e} inside {in_o[j]}/{in_c[j]} w {float(batch.priority_is[j]):.3f}')
