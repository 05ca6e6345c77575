"""Host <-> device copy ceiling of the end-to-end path.

N concurrent ranks (one per GPU, under torchrun like bench.py) do ONLY the pinned-memory traffic of the benchmark's
end-to-end step -- 315 MB host -> device and 619 MB device -> host per 16 384 signals -- with no kernel:

  * `pipeline`:   the chunks, streams and buffers of tebscat_scat1d_forward_host with the kernel left out
                  (tebscat_scat1d_host_copies_only);
  * `plain`:      one cudaMemcpyAsync per direction over the whole buffers on two streams (torch copy_), the
                  simplest possible upper bound;
  * `h2d`, `d2h`: each direction alone.

    python tools/host_copy_ceiling.py                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/host_copy_ceiling.py

Rank 0 prints one JSON line: aggregate and per-rank GB/s (max time over ranks), and the signals/s they would allow.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    sys.path.insert(0, p)

import torch                                   # noqa: E402
import torch.distributed as dist               # noqa: E402

J, Q, T, N = 6, 8, 64, 4800
N_SIG = 16384
STEPS = int(os.environ.get('CEILING_STEPS', '5'))


def main():
    from tebscat import Scattering1D, _lib
    from tebscat.sharding import max_over_ranks
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    S = Scattering1D(J, N, Q, T=T).to(dev)
    plan = S._plan_for(local)
    C, n_out = S._sched[1].n_paths, S._sched[1].n_out
    x_host = torch.randn(N_SIG, N).pin_memory()
    out_host = torch.empty(N_SIG, C, n_out).pin_memory()
    x_dev = torch.empty_like(x_host, device=dev)
    out_dev = torch.zeros(out_host.shape, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(STEPS):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0, dev) / STEPS

    def pipeline():
        _lib.check(lib.tebscat_scat1d_host_copies_only(plan.handle, x_host.data_ptr(), N_SIG, out_host.data_ptr()))

    def plain():
        with torch.cuda.stream(s1):
            x_dev.copy_(x_host, non_blocking=True)
        with torch.cuda.stream(s2):
            out_host.copy_(out_dev, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    def h2d():
        x_dev.copy_(x_host, non_blocking=True)
        torch.cuda.synchronize()

    def d2h():
        out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.synchronize()

    b_in, b_out = x_host.numel() * 4, out_host.numel() * 4
    res = {}
    for name, fn, nbytes in (('pipeline', pipeline, b_in + b_out), ('plain', plain, b_in + b_out),
                             ('h2d', h2d, b_in), ('d2h', d2h, b_out)):
        dt = timed(fn)
        res[name] = {'s_per_step': dt, 'gbs_per_rank': nbytes / dt / 1e9, 'gbs_aggregate': world * nbytes / dt / 1e9}
        if name in ('pipeline', 'plain'):
            res[name]['signals_per_s_allowed'] = world * N_SIG / dt
    if rank == 0:
        print(json.dumps({'tool': 'host_copy_ceiling', 'n_gpus': world, 'signals_per_step_per_gpu': N_SIG,
                          'h2d_bytes_per_step': b_in, 'd2h_bytes_per_step': b_out, 'steps': STEPS,
                          'host_cpus': os.cpu_count(), **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
