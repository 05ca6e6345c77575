"""BASELINE configs[3]: the scattering sweep J = 4..10, N = 2^12..2^16 (Q = 8, T = 2^J), batch-sharded over the GPUs
of one box with no collective -- the same launch contract as bench.py:

    python tools/sweep_config3.py                                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        tools/sweep_config3.py

Every point: synthetic randn batch sized to ~1 GB of output per GPU, capped so that a point stays within
a few seconds; device-resident timing with CUDA events, max over ranks; rank 0 prints one JSON line per point with
signals/s (whole job) and the reference-equivalent FFT TFLOP/s (SURVEY 8d flops per signal)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    sys.path.insert(0, p)

import torch                                   # noqa: E402
import torch.distributed as dist               # noqa: E402

POINTS = [(4, 2 ** 12), (6, 2 ** 12), (8, 2 ** 13), (8, 2 ** 14), (10, 2 ** 15), (10, 2 ** 16)]
MFLOP = {(4, 2 ** 12): 25.5, (6, 2 ** 12): 31.6, (8, 2 ** 13): 71.8, (8, 2 ** 14): 155.0, (10, 2 ** 15): 337.0,
         (10, 2 ** 16): 722.0}                 # reference-equivalent FFT MFLOP per signal
Q = 8


def main():
    from tebscat import Scattering1D
    from tebscat.sharding import max_over_ranks
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    only = os.environ.get('SWEEP_POINTS')
    pts = POINTS if not only else [POINTS[int(i)] for i in only.split(',')]
    for J, N in pts:
        T = 2 ** J
        S = Scattering1D(J, N, Q, T=T).to(dev)
        C, n_out = S.output_size(), S.ind_end[J] - S.ind_start[J]
        per_signal_out = C * n_out * 4
        B = max(64, min(int(1e9 // per_signal_out), (1 << 27) >> S.J_pad))     # ~1 GB of output, bounded workspace
        B = int(os.environ.get('SWEEP_BATCH', B))
        x = torch.randn(B, N, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
        for _ in range(3):                     # plan building, graph capture (the large level captures on the caller's
            out, _ = S(x)                      # buffers at their second sighting), steady state from the third call on
            del out
        torch.cuda.synchronize()
        reps = 3 if S.J_pad <= 14 else 2
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S(x)
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1), dev) / reps
        if rank == 0:
            rate = world * B / (ms * 1e-3)
            level = 'fused single kernel' if (S.J_pad <= 13 and not S._op_by_op) else 'large-support level'
            print(json.dumps({'tool': 'sweep_config3', 'J': J, 'Q': Q, 'T': T, 'N': N, 'J_pad': S.J_pad, 'paths': C,
                              'n_out': n_out, 'n_gpus': world, 'batch_per_gpu': B, 'ms': ms, 'signals_per_s': rate,
                              'ref_equiv_tflops': rate * MFLOP[(J, N)] * 1e6 / 1e12, 'level': level}), flush=True)
        del S, x
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
