"""Write the plan of a Scattering1D configuration to a file that ``tebscat_plan_load`` (include/tebscat.h) reads:

    python -m tebscat.export_plan --J 6 --shape 4800 --Q 8 --T 64 [--max-order 2] [--oversampling 0] out.tebplan

A consumer without Python then needs only libtebscat.so: tebscat_plan_load(path, device, &plan) and
tebscat_scat1d_forward(plan, x_dev, B, S_dev, stream).  The file also carries, for the consumer's convenience, the
output geometry in its description (n_paths = channels in meta()['key'] order, n_out = samples per channel)."""
import argparse
import ctypes
import os
import sys

import numpy as np

from . import _lib
from .schedule import build_plan


def plan_desc(plan):
    desc = _lib.PlanDesc()
    desc.abi_version = _lib.ABI_VERSION
    desc.N = plan.N
    desc.log2_Np = plan.geo.J_pad
    desc.pad_left = plan.geo.pad_left
    desc.n_paths = plan.n_paths
    desc.n_out = plan.n_out
    desc.n_threads = plan.n_threads
    desc.smem_complex = plan.smem_complex
    desc.n_tasks = plan.tasks.shape[0]
    desc.n_steps = plan.steps.shape[0]
    desc.border_mode = int(getattr(plan, 'border', 0))
    desc.scratch_complex = int(getattr(plan, 'scratch_complex', 0))
    return desc


def save_plan(plan, path):
    """Write a schedule (tebscat.schedule.build_plan and friends) as a plan file."""
    lib = _lib.load()
    desc = plan_desc(plan)
    arena = np.ascontiguousarray(plan.arena, np.float32)
    tasks = np.ascontiguousarray(plan.tasks, np.int32)
    steps = np.ascontiguousarray(plan.steps, np.int32)
    chan = np.ascontiguousarray(plan.chan, np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib.tebscat_plan_save(os.fsencode(path), ctypes.byref(desc), arena.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                               arena.size, tasks.ctypes.data_as(i32p), steps.ctypes.data_as(i32p), chan.ctypes.data_as(i32p),
                               chan.size)
    _lib.check(rc)
    return path


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--J', type=int, required=True)
    ap.add_argument('--shape', type=int, required=True)
    ap.add_argument('--Q', type=int, default=1)
    ap.add_argument('--T', type=int, default=None)
    ap.add_argument('--max-order', type=int, default=2)
    ap.add_argument('--oversampling', type=int, default=0)
    ap.add_argument('path')
    a = ap.parse_args(argv)
    plan = build_plan(a.J, a.shape, a.Q, a.T if a.T is not None else 2 ** a.J, a.max_order, oversampling=a.oversampling)
    save_plan(plan, a.path)
    print('%s: %d channels x %d samples, %d steps, %d filter floats' % (a.path, plan.n_paths, plan.n_out, plan.steps.shape[0],
                                                                          plan.arena.size))


if __name__ == '__main__':
    sys.exit(main())
