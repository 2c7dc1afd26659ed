#!/usr/bin/env python
"""Phase breakdown of the three row-tile kernels inside a running learner: runs the bench workload
for a few steps (CUDA graph), then reads the clock stamps CTA (0,0) took at its phase boundaries.
The per-layer accumulators (weight waits, layer_forward segments) are only compiled in with
`make EXTRA=-DASAC_PROBES` — they perturb the timing they measure (one global read-modify-write
per probe by thread 0) and cost 2.7 % of the step, so the product build leaves them out and
prints zeros for them."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

NAMES = {
    0: {0: 'setup+jobs', 1: 'pipe_init+stage x', 2: 'policy trunk+head', 3: 'per-row math', 4: 'critic inputs',
        5: 'target critic', 6: 'online critic (post)', 7: 'cluster combine', 31: 'v-trace, y'},
    1: {0: 'start', 1: 'setup+inputs', 2: 'critic forward', 3: 'head+loss', 4: 'head backward', 31: 'trunk backward'},
    2: {0: 'start', 1: 'setup+jobs', 2: 'policy forward', 3: 'sample+critic input', 4: 'critic forward+exchange',
        5: 'critic backward', 6: 'action-grad exchange', 7: 'policy loss', 31: 'policy backward'},
}


def main():
    sac, _ = bench.build_learner('cuda:0', seed=1, capacity=1 << 16, fill=1 << 16)
    buf = (C.c_int64 * 96)()
    for _ in range(20):
        sac.train()
    torch.cuda.synchronize()
    assert sac._lib.asac_debug_phase_clocks(buf) == 0  # also resets the weight-wait counters
    steps = 10
    for _ in range(steps):
        sac.train()
    torch.cuda.synchronize()
    assert sac._lib.asac_debug_phase_clocks(buf) == 0
    print(f'weight waits of CTA (0,0): {buf[29] / 1965.0 / steps:.2f} us per step over {buf[30] // steps} layer acquisitions')
    clk = np.array(buf[:], dtype=np.int64).reshape(3, 32)
    n_pass = max(int(clk[1, 24]), 1)
    print('layer_forward of CTA (0,0), per 16-row pass: ' + ', '.join(
        f'{name} {clk[1, 20 + i] / 1965.0 / n_pass:.2f} us' for i, name in
        enumerate(['K-split GEMM', 'barrier', 'epilogue', 'barrier'])) + f'  ({n_pass // steps} passes per step)')
    print(f'whole trunk layer (acquire + forward + release): {clk[1, 25] / 1965.0 / max(int(clk[1, 26]), 1):.2f} us; '
          f'target-critic phase of the last value pass: before trunk {(clk[0, 10] - clk[0, 4]) / 1965.0:.2f}, '
          f'trunk {(clk[0, 11] - clk[0, 10]) / 1965.0:.2f}, head + barrier {(clk[0, 12] - clk[0, 11]) / 1965.0:.2f} us')
    mhz = 1965.0
    print('k_q_backward setup detail (us): entry->jobs %.2f, stage_head %.2f, pipe_init %.2f, stage x %.2f' % tuple(
        (clk[1, j] - clk[1, i]) / mhz for i, j in ((0, 8), (8, 9), (9, 10), (10, 1))))
    for k, title in ((0, 'k_value_pass (last launch = post pass)'), (1, 'k_q_backward'), (2, 'k_policy_backward')):
        print(title)
        idx = [i for i in sorted(NAMES[k]) if clk[k, i] != 0 and i in NAMES[k]]
        for a, b in zip(idx[:-1], idx[1:]):
            print(f'  {NAMES[k][b]:28s} {(clk[k, b] - clk[k, a]) / mhz:7.2f} us')
        print(f'  {"total":28s} {(clk[k, idx[-1]] - clk[k, idx[0]]) / mhz:7.2f} us')
    sac.close()


if __name__ == '__main__':
    main()
