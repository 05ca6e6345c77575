"""A/B of the headline kernel: where the consumers of the signal's spectrum U0 find it (schedule.u0_in_scratch).
Variants: shared memory (round-1 layout); a per-CTA global scratch for every consumer (OP_STOREC after the root
transform, first-order multiplies through OP_GMULFOLD / OP_GMULFOLD2) under several cost-model assumptions -- they
change the schedule, not the arithmetic; 'split'.  For each: parity against the float64 oracle on a few signals,
bitwise agreement with the shared-memory variant, and the rate on B signals (CUDA events, 10 repetitions after 3
warm-ups)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from oracle.scattering1d_oracle import ScatteringOracle
from tebscat import Scattering1D
from tebscat import schedule as sch

CFGS = [(6, 4800, 8, 64, 2), (11, 5760, 4, 16, 1), (4, 4096, 8, 16, 2), (6, 4096, 8, 64, 2)]
VARIANTS = [('0', 64.0, True), ('1', 1e9, False), ('1', 1e9, True), ('1', 64.0, True), ('1', 128.0, True), ('split', 64.0, True),
            ('auto', 64.0, True)]
if os.environ.get('AB_SHORT'):                       # the two layouts only
    VARIANTS = [('0', float('inf'), True), ('1', float('inf'), True)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
only = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else range(len(CFGS))
for ci in only:
    J, N, Q, T, mo = CFGS[ci]
    x = torch.randn(B, N, generator=torch.Generator().manual_seed(1)).cuda()
    ref = ScatteringOracle(J, N, Q, T, mo)(x[:4].cpu().numpy())
    base = None
    for mode, bpc, pairs in VARIANTS:
        os.environ['TEBSCAT_U0_GLOBAL'] = mode
        sch.GSRC_BYTES_PER_CYCLE, sch.GSRC_PAIRS = bpc, pairs
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
        if base is None:
            base = out
        st = S._schedule().stats
        print('J=%d N=%d Q=%d T=%d order %d  U0 %-7s cost %6g B/cycle pairs %d: %3d steps, model %7d cycles, %.3f ms, %8.0f signals/s, '
              'worst path err %.2e, bitwise == shared: %s'
              % (J, N, Q, T, mo, getattr(S._schedule(), 'u0_mode', '?'), bpc, pairs, st['n_steps'], st['est_cycles'], ms, B / ms * 1e3, err,
                 bool(torch.equal(out, base))), flush=True)
        del S
