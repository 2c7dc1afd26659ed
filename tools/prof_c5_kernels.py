"""Kernel-level table of one config-5 step (conv + attention representation through the torch-module bridge):
torch.profiler over eager steps (ASAC_BRIDGE_GRAPH=0), top kernels by device time."""
import os, sys
os.environ['ASAC_BRIDGE_GRAPH'] = '0'
sys.path[:0] = ['/root/repo', '/root/repo/advanced-soft-actor-critic_b200']
import torch, bench
from torch.profiler import profile, ProfilerActivity
bench.CFG.update(bench.CONFIGS['c5'])
sac, rng = bench.build_learner('cuda:0', seed=1, capacity=8192, fill=None)
for _ in range(5): sac.train()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): sac.train()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in evs)
print(f'device time per step: {tot / 3 / 1e3:.3f} ms over {sum(e.count for e in evs) // 3} kernel launches')
for e in sorted(evs, key=lambda e: -e.device_time_total)[:28]:
    print(f'{e.device_time_total / 3:9.1f} us  {e.count // 3:4d}x  {e.key[:110]}')
