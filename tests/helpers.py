"""Shared helpers of the test-suite (oracle access, emulator binding, tolerances)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
EMU_PATH = os.path.join(ROOT, 'tests', 'emu', 'libemu_scat.so')

CONFIGS = {
    # name: (J, N, Q, T, max_order)
    'H': (6, 4800, 8, 64, 2),
    'P': (11, 5760, 4, 16, 1),
    'S': (4, 1000, 4, 16, 2),
    'T': (5, 700, 2, 8, 2),
    'K': (6, 512, 16, 64, 2),
    'O': (5, 1200, 4, 32, 2),           # golden fixture with oversampling = 1
}
OVERSAMPLING = {'O': 1}


def rel_l2(a, b, axis=None):
    return np.linalg.norm(a - b, axis=axis) / np.maximum(np.linalg.norm(b, axis=axis), 1e-30)


def path_tolerance(ref64, ref32, rel=1e-5, noise_factor=4.0):
    """Per-path L2 error budget: rel-L2 <= 1e-5 (north_star), or -- on paths whose energy
    is so far below the signal's that the REFERENCE's own fp32 arithmetic is noisier than
    that -- `noise_factor` times the single-precision oracle's distance to the float64 one."""
    return np.maximum(rel * np.linalg.norm(ref64, axis=-1),
                      noise_factor * np.linalg.norm(ref32.astype(np.float64) - ref64, axis=-1))


def phase_path_tolerance(oracle_aligned_to_ref, ref_fp32, rel=1e-5, noise_factor=4.0):
    """Per-path L2 error budget of the phase path: rel-L2 <= 1e-5 (north_star), or -- on paths where the
    REFERENCE's own fp32 arithmetic (p * theta rounded in fp32 with p up to ~100) is noisier
    than that -- `noise_factor` times the distance of the live reference's committed fp32 output to the
    branch-aligned float64 oracle.  On randn rows the reference sits below 1e-5 on every path, so the
    budget there is the plain north-star bound times at most `noise_factor`; the tests assert the plain
    bound on those rows separately."""
    return np.maximum(rel * np.linalg.norm(oracle_aligned_to_ref, axis=-1),
                      noise_factor * np.linalg.norm(ref_fp32.astype(np.float64) - oracle_aligned_to_ref, axis=-1))


def emu_available():
    return os.path.exists(EMU_PATH)


def emu_forward(plan, x, stage_a=False, epilogue=None):
    """Run the schedule of `plan` through the host emulator (tests/emu).  With stage_a=True the
    plan is a phase stage-A schedule and the (cartesian, polar) analytic signals are returned."""
    lib = ctypes.CDLL(EMU_PATH)
    x = np.ascontiguousarray(x, np.float32)
    B = x.shape[0]
    ep_mean = ep_std = ep_mode = None
    ep_eps, ep_trim, ep_tm = 0.0, 0, 0
    oshape = (B, plan.n_paths, plan.n_out)
    if epilogue is not None:                      # dict(mean, std, mode, log_eps, trim, time_major): SURVEY 8f-2
        ep_mean = np.ascontiguousarray(epilogue['mean'], np.float32)
        ep_std = np.ascontiguousarray(epilogue['std'], np.float32)
        ep_mode = np.ascontiguousarray(epilogue['mode'], np.uint8)
        ep_eps, ep_trim, ep_tm = float(epilogue['log_eps']), int(epilogue['trim']), int(bool(epilogue['time_major']))
        keep = plan.n_out - 2 * ep_trim
        oshape = (B, keep, plan.n_paths) if ep_tm else (B, plan.n_paths, keep)
    out = np.full(oshape if not stage_a else (1,), np.nan, np.float32)
    zshape = (B, plan.n_paths, plan.n_out, 2) if stage_a else (1,)
    zc, zp = np.full(zshape, np.nan, np.float32), np.full(zshape, np.nan, np.float32)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    tasks = np.ascontiguousarray(plan.tasks, np.int32)
    steps = np.ascontiguousarray(plan.steps, np.int32)
    arena = np.ascontiguousarray(plan.arena, np.float32)
    chan = np.ascontiguousarray(plan.chan, np.int32)
    rc = lib.emu_scat1d_forward(plan.N, plan.geo.J_pad, plan.geo.pad_left, plan.n_paths, plan.n_out,
                                plan.smem_complex, tasks.shape[0], steps.shape[0],
                                arena.ctypes.data_as(fp), tasks.ctypes.data_as(ip), steps.ctypes.data_as(ip),
                                chan.ctypes.data_as(ip), x.ctypes.data_as(fp), ctypes.c_longlong(B), out.ctypes.data_as(fp),
                                zc.ctypes.data_as(fp), zp.ctypes.data_as(fp), 3 if stage_a else 0,
                                ep_mean.ctypes.data_as(fp) if ep_mean is not None else None,
                                ep_std.ctypes.data_as(fp) if ep_std is not None else None,
                                ep_mode.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)) if ep_mode is not None else None,
                                ctypes.c_float(ep_eps), ep_trim, ep_tm, int(getattr(plan, 'border', 0)),
                                int(getattr(plan, 'scratch_complex', 0)))
    assert rc == 0
    if stage_a:
        return zc[..., 0] + 1j * zc[..., 1], zp
    return out


def emu_forward_gsrc(plan, src, out, chan_shift=0):
    """Fused subtree of the large-support level through the host emulator: `src` (B, L) complex spectra in
    bit-reversed order -> the plan's channels of `out` (B, n_paths, n_out) float32, shifted by `chan_shift`
    channels (the C entry point takes the shifted pointer, tebscat_scat1d_forward_gsrc)."""
    lib = ctypes.CDLL(EMU_PATH)
    src = np.ascontiguousarray(src, np.complex64)
    B, L = src.shape
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (B, plan.n_paths, plan.n_out)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    tasks = np.ascontiguousarray(plan.tasks, np.int32)
    steps = np.ascontiguousarray(plan.steps, np.int32)
    arena = np.ascontiguousarray(plan.arena, np.float32)
    chan = np.ascontiguousarray(plan.chan, np.int32)
    lib.emu_scat1d_forward_gsrc.argtypes = [ctypes.c_int] * 5 + [fp, ip, ip, ip, ctypes.c_void_p, ctypes.c_longlong,
                                                                 ctypes.c_longlong, ctypes.c_void_p]
    rc = lib.emu_scat1d_forward_gsrc(plan.n_paths, plan.n_out, plan.smem_complex, tasks.shape[0], steps.shape[0],
                                     arena.ctypes.data_as(fp), tasks.ctypes.data_as(ip), steps.ctypes.data_as(ip),
                                     chan.ctypes.data_as(ip), ctypes.c_void_p(src.ctypes.data), L, B,
                                     ctypes.c_void_p(out.ctypes.data + 4 * chan_shift * plan.n_out))
    assert rc == 0
    return out


def emu_pair_stage(pair_plan, zp_rows, zc_rows, powers):
    """Phase stage B in its transform form through the host emulator: rows of (|z_i|, theta_i) and
    (re, im) of z_j, one power per row -> (rows, n_out) low-passed real parts."""
    lib = ctypes.CDLL(EMU_PATH)
    rows, N = zp_rows.shape[0], zp_rows.shape[1]
    rpj = pair_plan.n_paths
    jobs = -(-rows // rpj)
    pad = jobs * rpj - rows

    def padded(a):
        a = np.ascontiguousarray(a, np.float32)
        return np.ascontiguousarray(np.concatenate([a, np.repeat(a[-1:], pad, axis=0)]) if pad else a)
    zp, zc, pw = padded(zp_rows), padded(zc_rows), padded(np.asarray(powers, np.float32))
    out = np.full((jobs, rpj, pair_plan.n_out), np.nan, np.float32)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    tasks = np.ascontiguousarray(pair_plan.tasks, np.int32)
    steps = np.ascontiguousarray(pair_plan.steps, np.int32)
    arena = np.ascontiguousarray(pair_plan.arena, np.float32)
    chan = np.ascontiguousarray(pair_plan.chan, np.int32)
    rc = lib.emu_scat1d_forward(pair_plan.N, pair_plan.geo.J_pad, pair_plan.geo.pad_left, rpj, pair_plan.n_out,
                                pair_plan.smem_complex, tasks.shape[0], steps.shape[0],
                                arena.ctypes.data_as(fp), tasks.ctypes.data_as(ip), steps.ctypes.data_as(ip),
                                chan.ctypes.data_as(ip), pw.ctypes.data_as(fp), ctypes.c_longlong(jobs), out.ctypes.data_as(fp),
                                zc.ctypes.data_as(fp), zp.ctypes.data_as(fp), 4, None, None, None, ctypes.c_float(0.0), 0, 0,
                                int(getattr(pair_plan, 'border', 0)), 0)
    assert rc == 0
    return out.reshape(jobs * rpj, pair_plan.n_out)[:rows]
