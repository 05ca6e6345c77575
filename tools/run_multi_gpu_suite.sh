#!/bin/bash
# One 8-GPU box: host-copy ceiling and configs[3] sweep at 8/4/2 ranks, bench.py at 8 ranks.  Outputs in gpurun_out/.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export CEILING_STEPS=3
for n in 8 4 2 1; do
  timeout 120 $TR --nproc-per-node $n --master-port $((29600+n)) tools/host_copy_ceiling.py > gpurun_out/r2_ceiling_n$n.json 2> gpurun_out/r2_ceiling_n$n.err
done
timeout 200 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 5 --warmup 3 --no-phase > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
for n in 8 4 2; do
  timeout 240 $TR --nproc-per-node $n --master-port $((29800+n)) tools/sweep_config3.py > gpurun_out/r2_sweep3_n$n.jsonl 2> gpurun_out/r2_sweep3_n$n.err
done
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
lscpu | head -25 > gpurun_out/r2_lscpu.txt 2>&1
numactl -H >> gpurun_out/r2_lscpu.txt 2>&1 || true
tail -2 gpurun_out/r2_sweep3_n8.jsonl; cat gpurun_out/r2_ceiling_n8.json | cut -c1-600
