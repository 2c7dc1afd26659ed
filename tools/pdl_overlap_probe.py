#!/usr/bin/env python
"""%globaltimer stamps around the critics' Adam step inside a running learner (library built with
`make EXTRA=-DASAC_PROBES`): does the policy backward start while the optimiser kernel runs (programmatic dependent
launch), and how long does it wait for it?"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'advanced-soft-actor-critic_b200')]
import torch  # noqa: E402

import bench  # noqa: E402

sac, _ = bench.build_learner('cuda:0', seed=1, capacity=1 << 16, fill=1 << 16)
buf = (C.c_uint64 * 8)()
for _ in range(20):
    sac.train()
torch.cuda.synchronize()
sac._lib.asac_debug_global_stamps(buf)
for _ in range(5):
    sac.train()
    torch.cuda.synchronize()
    sac._lib.asac_debug_global_stamps(buf)
    t = list(buf)
    base = t[0]
    names = {0: 'critic backward exit (last CTA)', 4: 'policy backward first CTA start',
             5: 'policy backward last CTA start', 6: 'policy backward last wait return'}
    print(' | '.join(f'{n}: {(int(t[i]) - int(base)) / 1e3:+.2f} us' for i, n in names.items()))
sac.close()
