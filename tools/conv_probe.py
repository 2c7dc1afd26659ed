"""cuDNN / native convolution timings for the reference's 'simple' visual stack (tests/nn_conv_attn.py)."""
import sys, time
sys.path[:0] = ['/root/repo', '/root/repo/advanced-soft-actor-critic_b200']
import torch
import asac_b200.nn_models as m
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
x = torch.randn(512 * 14, 3, 30, 30, device='cuda')
conv = m.ConvLayers(30, 30, 3, 'simple', out_dense_depth=2, output_size=8).cuda()
def fb():
    out = conv.conv_layers(x); out.sum().backward()
for enabled in (True, False):
    for tf32 in (True, False):
        for bench in (False, True):
            with torch.backends.cudnn.flags(enabled=enabled, allow_tf32=tf32, benchmark=bench):
                with torch.no_grad():
                    f = t(lambda: conv.conv_layers(x))
                print(f'cudnn={enabled} tf32={tf32} benchmark={bench}: fwd {f:.2f} ms, fwd+bwd {t(fb):.2f} ms', flush=True)
xc = x.contiguous(memory_format=torch.channels_last)
conv = conv.to(memory_format=torch.channels_last)
with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
    print('channels_last fp32 fwd %.2f ms' % t(lambda: conv.conv_layers(xc)), flush=True)
