#!/usr/bin/env python
"""Where the host time of the end-to-end loop goes (put_episode + train() + D2H per step): cProfile over the loop
bench.py times as `e2e`."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import torch  # noqa: E402

import bench  # noqa: E402

sac, rng = bench.build_learner('cuda:0', seed=1, capacity=1 << 16, fill=1 << 16)
S, A = bench.CFG['obs_shape'][0], bench.CFG['A']
eps = [bench.pin_episode(bench.synth_episode(rng, 8, S, A)) for _ in range(32)]
td_hosts = [torch.empty(sac.batch_size, dtype=torch.float32).pin_memory() for _ in range(2)]
td_events = [torch.cuda.Event() for _ in range(2)]
for i in range(50):
    sac.put_episode(**eps[i % 32]); sac.train()
torch.cuda.synchronize()


def loop(steps):
    acc = 0.0
    for i in range(steps):
        sac.put_episode(**eps[i % 32])
        sac.train()
        td_hosts[i & 1].copy_(sac._wk['td_error'], non_blocking=True)
        td_events[i & 1].record()
        if i > 0:
            td_events[(i - 1) & 1].synchronize()
            acc += float(td_hosts[(i - 1) & 1][0])
    torch.cuda.synchronize()
    return acc


t0 = time.perf_counter(); loop(2000); dt = time.perf_counter() - t0
print(f'e2e loop: {dt / 2000 * 1e6:.1f} us per step')
# host-only cost: the same calls without waiting for results
t0 = time.perf_counter()
for i in range(2000):
    sac.put_episode(**eps[i % 32])
t1 = time.perf_counter()
torch.cuda.synchronize()
for i in range(2000):
    sac.train()
t2 = time.perf_counter()
torch.cuda.synchronize()
print(f'host: put_episode {(t1 - t0) / 2000 * 1e6:.1f} us, train() enqueue {(t2 - t1) / 2000 * 1e6:.1f} us per call')
pr = cProfile.Profile(); pr.enable(); loop(2000); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
sac.close()
