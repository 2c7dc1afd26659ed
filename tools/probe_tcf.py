"""Phase clocks of the features-on-M tcgen05 layer engine (asac_mlp_forward_tcf_probe): where a layer's time goes
for row tiles of 8 .. 256 rows.  Prints cycles per phase of CTA 0 (SM clock)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import torch
from asac_b200 import _lib, lowering
from asac_b200._lib import check, ptr

lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
in_dim, H, depth, O = 8, 64, 3, 1
shape = lowering.NetShape(in_dim, H, depth, O)
flat = (torch.rand(shape.stride, device='cuda') - 0.5) * 0.3
for rc in (8, 16, 32, 64, 208):
    for variant in (0, 1):
        x = torch.randn(rc, in_dim, device='cuda')
        out = torch.zeros(rc, O, device='cuda')
        probe = torch.zeros((depth + 1) * 5, dtype=torch.int64, device='cuda')
        for _ in range(3):
            check(lib.asac_mlp_forward_tcf_probe(ptr(flat), in_dim, H, depth, O, ptr(x), rc, ptr(out), rc, variant,
                                                 ptr(probe), s), 'probe')
        torch.cuda.synchronize()
        p = probe.cpu().view(depth + 1, 5)
        rows = []
        for l in range(depth + 1):
            t = p[l].tolist()
            rows.append((t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3]) if variant else (t[4] - t[0],))
        print(f'rows {rc:3d} variant {variant}: per layer (issue, wait-for-MMA, epilogue, store+sync) cycles:', rows,
              'layer-to-layer', [int(p[l + 1][0] - p[l][0]) for l in range(depth)])
    # whole-kernel time against the FFMA kernel, one CTA
    for name, fn in (('tcf', lambda: lib.asac_mlp_forward_tcf(ptr(flat), in_dim, H, depth, O, ptr(x), rc, ptr(out), rc, 0, s)),
                     ('ffma', lambda: lib.asac_mlp_forward(ptr(flat), in_dim, H, depth, O, ptr(x), rc, ptr(out), s))):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f'   {name}: {e0.elapsed_time(e1) * 5:.2f} us per launch ({rc} rows, one CTA)')
