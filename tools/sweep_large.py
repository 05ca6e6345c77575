"""BASELINE configs[3]: scattering sweep J=4..10, N=2^12..2^16 (Q=8, T=2^J) -- signals/s on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import Scattering1D
SWEEP = [(4, 2 ** 12), (6, 2 ** 12), (8, 2 ** 13), (8, 2 ** 14), (10, 2 ** 15), (10, 2 ** 16)]
FLOPS = {(4, 2 ** 12): 25.5e6, (6, 2 ** 12): 31.6e6, (8, 2 ** 13): 71.8e6, (8, 2 ** 14): 155e6, (10, 2 ** 15): 337e6, (10, 2 ** 16): 722e6}
for J, N in SWEEP:
    S = Scattering1D(J, N, 8, T=2 ** J).cuda()
    C, n_out = S.output_size(), None
    B = max(8, min(4096, int(1e9 / (C * max(1, N >> J) * 4))))       # ~1 GB of output, capped
    B = min(B, 2048 if S.J_pad <= 13 else max(16, (1 << 27) >> S.J_pad))
    x = torch.randn(B, N, device='cuda')
    out, _ = S(x); torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): S(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = B / ms * 1e3
    print('J=%2d N=2^%d Np=2^%d paths=%d out=%s B=%d: %.1f ms -> %.0f signals/s, %.1f TFLOP/s ref-equivalent'
          % (J, N.bit_length() - 1, S.J_pad, out.shape[1], tuple(out.shape[1:]), B, ms, rate, rate * FLOPS[(J, N)] / 1e12), flush=True)
