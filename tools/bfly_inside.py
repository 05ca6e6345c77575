"""Phases of one forward radix-16 butterfly INSIDE the interpreter (library built with
-DTEBSCAT_PROF_PHASES -DTEBSCAT_PROF_BFLY): setup+loads / dft16 / twiddle+stores, warp 0 of CTA 0."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import numpy as np, torch, copy
from tebscat import _lib
from tebscat.schedule import build_plan, OP_NOP, OP_FFT, OP_LOAD, TASK_INTS
from tebscat.torch_frontend import _DevicePlan
base = build_plan(6, 4800, 8, 64, 2)
lib = _lib.load()
lib.tebscat_debug_bfly.argtypes = [ctypes.c_void_p, ctypes.c_int]

def run(tasks_per_step, label):
    p = copy.copy(base)
    rows, ranges = [], []
    for st in tasks_per_step:
        ranges.append([len(rows), len(rows) + len(st)]); rows += st
    p.tasks = np.asarray(rows, np.int32).reshape(-1, TASK_INTS)
    p.steps = np.asarray(ranges, np.int32)
    dp = _DevicePlan(p, 0)
    B = 148
    x = torch.randn(B, 4800, device='cuda'); out = torch.zeros(B, p.n_paths, p.n_out, device='cuda')
    ns = len(ranges)
    clk = np.zeros(4 * ns + 1, np.int64)
    _lib.check(lib.tebscat_scat1d_profile_steps(dp.handle, x.data_ptr(), B, out.data_ptr(), clk.ctypes.data, torch.cuda.current_stream().cuda_stream))
    lib.tebscat_debug_bfly(None, 1)
    _lib.check(lib.tebscat_scat1d_profile_steps(dp.handle, x.data_ptr(), B, out.data_ptr(), clk.ctypes.data, torch.cuda.current_stream().cuda_stream))
    d = np.zeros(8, np.int64); lib.tebscat_debug_bfly(d.ctypes.data, 0)
    start = clk[:ns]; ph = clk[ns + 1:].reshape(ns, 3)
    r = np.stack([ph[:, 0] - start, ph[:, 1] - ph[:, 0], ph[:, 2] - ph[:, 1], clk[1:ns + 1] - start], 1)[4:]
    n = max(1, d[3])
    print('   last step: decode-end -> fft_task %d -> fft_pass %d -> butterfly start %d cycles' % (d[4] - d[7], d[5] - d[4], d[6] - d[5]))
    print('%-40s step %5d = decode %4d + body %5d + bar-issue %4d + rest %4d | butterfly: setup+loads %4d dft %4d twiddle+store %4d' % (
        label, np.median(r[:, 3]), np.median(r[:, 0]), np.median(r[:, 1]), np.median(r[:, 2]),
        np.median(r[:, 3] - r[:, 0] - r[:, 1] - r[:, 2]), d[0] // n, d[1] // n, d[2] // n))

load = [OP_LOAD, 0, 512, 0, 0, 0, 0, 0, 0, 0, 0, 0]
f16 = lambda nb, nt, logB: [OP_FFT, 0, nt, 0, nb, logB, 4, 0, 0, 0, 0, 0]
run([[load]] + [[f16(32, 32, 9)]] * 64, 'one warp, one R16 butterfly (s=32)')
run([[load]] + [[f16(128, 128, 11)]] * 64, '4 warps, one R16 butterfly each')
run([[load]] + [[f16(256, 256, 12)]] * 64, '8 warps')
run([[load]] + [[f16(512, 512, 13)]] * 64, '16 warps, one R16 butterfly each (s=512)')
run([[load]] + [[f16(1024, 512, 13)]] * 64, '16 warps, two R16 butterflies each')
