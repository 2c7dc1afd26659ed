"""Where a config-5 step (conv + attention representation through the torch-module bridge) spends its time:
wall-clock per phase with a synchronize after each (no profiler)."""
import sys, time
sys.path[:0] = ['/root/repo', '/root/repo/advanced-soft-actor-critic_b200']
import torch, bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bench.CFG.update(bench.CONFIGS['c5']); bench.CFG['B'] = B
sac, rng = bench.build_learner('cuda:0', seed=1, capacity=8192, fill=None)
for _ in range(3): sac.train()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): sac.train()
torch.cuda.synchronize(); print('B', B, 'ms/step', (time.perf_counter() - t0) / 3 * 1e3, flush=True)
br = sac._bridge
orig = br.l_states
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); print(f'   {name}: {(time.perf_counter() - t) * 1e3:.2f} ms', flush=True)
        return r
    return w
br.l_states = timed('l_states', orig)
br.backward = timed('backward', br.backward)
sac._enqueue_sample = timed('sample+gather', sac._enqueue_sample)
sac.train(); torch.cuda.synchronize()
# inside the representation: conv vs attention
st = sac._sets[0]; bt = st['bt']
obs = sac._process_torch_obs_list(list(bt['obs_list']))
m = sac.model_rep
with torch.no_grad():
    for name, fn in (('conv', lambda: m.conv(obs[1])),):
        torch.cuda.synchronize(); t = time.perf_counter(); v = fn(); torch.cuda.synchronize()
        print(f'   {name}: {(time.perf_counter() - t) * 1e3:.2f} ms', tuple(v.shape), flush=True)
