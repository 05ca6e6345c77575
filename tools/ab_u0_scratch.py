"""A/B of the headline kernel: U0 in shared memory (default) against U0 parked in a per-CTA global scratch
(TEBSCAT_U0_GLOBAL=1: OP_STOREC after the root transform, first-order multiplies through OP_GMULFOLD).  For each
mode: parity against the float64 oracle on a few signals, bitwise agreement between the modes, and the rate on
16384 signals (CUDA events, 10 repetitions after 3 warm-ups)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from oracle.scattering1d_oracle import ScatteringOracle
from tebscat import Scattering1D
from tebscat.synth import ctg_batch

CFGS = [(6, 4800, 8, 64, 2), (11, 5760, 4, 16, 1), (4, 4096, 8, 16, 2)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for J, N, Q, T, mo in CFGS:
    x = torch.randn(B, N, generator=torch.Generator().manual_seed(1)).cuda()
    ref = ScatteringOracle(J, N, Q, T, mo)(x[:4].cpu().numpy())
    outs = {}
    for mode in ('0', '1'):
        os.environ['TEBSCAT_U0_GLOBAL'] = mode
        S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
        out = S(x)[0]
        torch.cuda.synchronize()
        o4 = out[:4].cpu().numpy().astype(np.float64)
        err = (np.linalg.norm(o4 - ref, axis=-1) / np.linalg.norm(ref, axis=-1)).max()
        for _ in range(3):
            S(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            S(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        outs[mode] = out
        print('J=%d N=%d Q=%d T=%d order %d  U0 %s: %d steps, %.3f ms, %.0f signals/s, worst path err %.2e'
              % (J, N, Q, T, mo, 'in global scratch' if mode == '1' else 'in shared memory ', S._schedule().stats['n_steps'],
                 ms, B / ms * 1e3, err), flush=True)
    print('   modes bitwise equal:', bool(torch.equal(outs['0'], outs['1'])), flush=True)
