"""Measure the fixed per-step overhead of the interpreter with synthetic schedules."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch
from tebscat import _lib
from tebscat.schedule import build_plan, OP_NOP, OP_FFT, OP_LOAD, TASK_INTS
from tebscat.torch_frontend import _DevicePlan

base = build_plan(6, 4800, 8, 64, 2)

def run(tasks_per_step, label):
    import copy
    p = copy.copy(base)
    rows, ranges = [], []
    for st in tasks_per_step:
        ranges.append([len(rows), len(rows) + len(st)]); rows += st
    p.tasks = np.asarray(rows, np.int32).reshape(-1, TASK_INTS)
    p.steps = np.asarray(ranges, np.int32)
    dp = _DevicePlan(p, 0)
    B = 148
    x = torch.randn(B, 4800, device='cuda'); out = torch.zeros(B, p.n_paths, p.n_out, device='cuda')
    clk = np.zeros(len(ranges) + 1, np.int64); best = None
    for _ in range(5):
        _lib.check(_lib.load().tebscat_scat1d_profile_steps(dp.handle, x.data_ptr(), B, out.data_ptr(), clk.ctypes.data,
                                                            torch.cuda.current_stream().cuda_stream))
        d = np.diff(clk); best = d if best is None else np.minimum(best, d)
    print('%-44s median %6d  min %6d cycles/step' % (label, np.median(best[2:]), best[2:].min()))

nop = [OP_NOP, 0, 512, 0, 0, 0, 0, 0, 0, 0, 0, 0]
run([[nop]] * 64, 'NOP task on all warps')
load = [OP_LOAD, 0, 512, 0, 0, 0, 0, 0, 0, 0, 0, 0]
f16 = lambda nb, nt, logB: [OP_FFT, 0, nt, 0, nb, logB, 4, 0, 0, 0, 0, 0]
run([[load]] + [[f16(32, 32, 9)]] * 64, 'one warp, one R16 butterfly (s=32)')
run([[load]] + [[f16(512, 512, 13)]] * 64, '16 warps, one R16 butterfly each (s=512)')
run([[load]] + [[f16(512, 512, 4)]] * 64, '16 warps, R16 unit stride (no twiddles)')
run([[load]] + [[f16(256, 256, 12)]] * 64, '8 warps, one R16 butterfly each (s=256)')
f8 = lambda nb, nt, logB: [OP_FFT, 0, nt, 0, nb, logB, 3, 0, 0, 0, 0, 0]
run([[load]] + [[f8(1024, 512, 13)]] * 64, '16 warps, two R8 butterflies each (s=1024)')
run([[load]] + [[f8(512, 512, 3)]] * 64, '16 warps, one R8 unit stride')

# ---- MULFOLD variants (filters: offset 0 of the arena = phi level 0, 8192 floats) -------------
OP_MULFOLD = 3
mf = lambda logk, nt, mask, sexp=13: [OP_MULFOLD | (sexp << 8), 0, nt, 0, 13, logk, 8192, 0, mask, 0, 2 if logk >= 2 else 0, 0]   # filter offset 0: any data
run([[load]] + [[mf(0, 512, 0)]] * 32, 'MULFOLD k=1  8192->8192, 512 thr (4 items/thr)')
run([[load]] + [[mf(1, 512, 0)]] * 32, 'MULFOLD k=2  8192->4096, 512 thr (4 items/thr)')
run([[load]] + [[mf(2, 512, 1)]] * 32, 'MULFOLD k=4  1 chunk, 2048 out, 512 thr (1 trip)')
run([[load]] + [[mf(2, 256, 1)]] * 32, 'MULFOLD k=4  1 chunk, 2048 out, 256 thr (2 trips)')
run([[load]] + [[mf(6, 32, 0x8001)]] * 32, 'MULFOLD k=64 2 chunks, 128 out, 32 thr (1 trip)')
run([[load]] + [[mf(4, 128, 0xf)]] * 32, 'MULFOLD k=16 4 chunks, 512 out, 128 thr (1 trip)')
