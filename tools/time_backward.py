"""Time the backward pass of the headline configuration: forward + backward per signal."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'vae-teb_b200'))
from tebscat import Scattering1D, _lib   # noqa: E402
from tebscat.synth import ctg_batch      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = Scattering1D(6, 4800, 8, T=64).cuda()
x = ctg_batch(B // 2, 4800, seed=1).reshape(B, 4800).cuda().requires_grad_(True)
out, _ = S(x)
w = torch.randn_like(out)
for it in range(3):
    x.grad = None
    out, _ = S(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out.backward(w)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('B=%d backward %.2f ms (device) %.2f ms (host wall), %d launches -> %.0f signals/s' % (
        B, e0.elapsed_time(e1), 1e3 * (t1 - t0), _lib.load().tebscat_last_launch_count(), B / (e0.elapsed_time(e1) * 1e-3)))
