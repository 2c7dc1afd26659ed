"""Error of the three forward engines against a float64 forward of the same net: FFMA row tiles (asac_mlp_forward),
tcgen05 3xTF32 with one accumulator (variant 0) and with the cross terms in their own accumulator (variant 2)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import numpy as np
import torch
from asac_b200 import _lib, lowering
from asac_b200._lib import check, ptr
from oracle.sac_oracle import init_q, q_forward

lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
gen = torch.Generator().manual_seed(5)
for (in_dim, depth, rows, rc) in [(8, 3, 4096, 16), (8, 3, 4096, 96), (8, 3, 4160, 208), (8, 2, 4096, 32)]:
    S, A, H = in_dim - 2, 2, 64
    p = init_q(S, A, H, depth, gen)
    x = torch.randn(rows, in_dim, generator=gen)
    p64 = {k: v.double() for k, v in p.items()}
    ref64 = q_forward(p64, depth, x[:, :S].double(), x[:, S:].double()).numpy().reshape(-1)
    ref32 = q_forward(p, depth, x[:, :S], x[:, S:]).numpy().reshape(-1)
    flat = lowering.flat_from_state_dict(lowering.NetShape(in_dim, H, depth, 1), p, policy=False).cuda()
    xc = x.cuda().contiguous()
    scale = np.abs(ref64).max()
    res = {'torch32': ref32}
    out = torch.zeros(rows, 1, device='cuda')
    check(lib.asac_mlp_forward(ptr(flat), in_dim, H, depth, 1, ptr(xc), rows, ptr(out), s), 'ffma')
    res['ffma'] = out.cpu().numpy().reshape(-1).copy()
    for v in (0, 2):
        if v == 2 and 2 * rc > 512:
            continue
        out.zero_()
        check(lib.asac_mlp_forward_tcf(ptr(flat), in_dim, H, depth, 1, ptr(xc), rows, ptr(out), rc, v, s), 'tcf')
        res[f'tcf_v{v}'] = out.cpu().numpy().reshape(-1).copy()
    print(f'in {in_dim} depth {depth} rows/cta {rc}:', {k: f'max {np.abs(v - ref64).max() / scale:.2e} rms {np.sqrt(np.mean((v - ref64) ** 2)) / scale:.2e} mean {np.mean(v - ref64) / scale:+.1e}' for k, v in res.items()})
