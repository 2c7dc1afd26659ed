#!/usr/bin/env python
"""Throughput of the stock-net forward kernels on a large row count: the exact-fp32 FFMA row-tile
kernel (asac_mlp_forward) and the tcgen05 3xTF32 kernel (asac_mlp_forward_tc).

    python tools/bench_mlp_forward.py [--rows N] [--iters K] [--net q|policy]

Prints one JSON line per kernel: rows/s, algorithmic TFLOP/s (2 * MACs of the fp32 net; the
3xTF32 kernel issues 3x that on the tensor pipe, reported as `tensor_tflops_issued`)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import torch  # noqa: E402

from asac_b200 import _lib, lowering  # noqa: E402
from asac_b200._lib import check, ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=1 << 20)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--net', default='q')
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    lib = _lib.load()
    torch.cuda.set_device(0)
    in_dim, H, depth, out = (8, 64, 3, 1) if args.net == 'q' else (6, 64, 3, 4)
    shape = lowering.NetShape(in_dim, H, depth, out)
    gen = torch.Generator(device='cuda').manual_seed(0)
    flat = (torch.rand(shape.stride, device='cuda', generator=gen) - 0.5) * 0.3
    x = torch.randn(args.rows, in_dim, device='cuda', generator=gen)
    y = torch.zeros(args.rows, out, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    flops_row = 2 * (in_dim * H + (depth - 1) * H * H + H * out)
    issued_row = 3 * 2 * ((-(-in_dim // 8) * 8) * H + (depth - 1) * H * H + H * 16)
    for name, fn in (('ffma_fp32', lib.asac_mlp_forward), ('tcgen05_3xtf32', lib.asac_mlp_forward_tc)):
        if args.only and args.only != name:
            continue
        for _ in range(3):
            check(fn(ptr(flat), in_dim, H, depth, out, ptr(x), args.rows, ptr(y), s), name)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            check(fn(ptr(flat), in_dim, H, depth, out, ptr(x), args.rows, ptr(y), s), name)
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1) / args.iters
        print(json.dumps({'kernel': name, 'net': args.net, 'rows': args.rows, 'ms': ms,
                          'rows_per_s': args.rows / (ms * 1e-3),
                          'algorithmic_tflops': flops_row * args.rows / (ms * 1e-3) / 1e12,
                          'tensor_tflops_issued': (issued_row * args.rows / (ms * 1e-3) / 1e12) if 'tc' in name else None,
                          'hbm_GBps': args.rows * 4 * (in_dim + out) / (ms * 1e-3) / 1e9}))


if __name__ == '__main__':
    main()
