#!/bin/bash
# configs[3] sweep at 8, 4 and 2 ranks on one 8-GPU box (large-support points only at 4 and 2 to bound the time)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29808 tools/sweep_config3.py > gpurun_out/r2b_sweep3_n8.jsonl 2> gpurun_out/r2b_sweep3_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29804 tools/sweep_config3.py > gpurun_out/r2b_sweep3_n4.jsonl 2> gpurun_out/r2b_sweep3_n4.err
timeout 300 $TR --nproc-per-node 2 --master-port 29802 tools/sweep_config3.py > gpurun_out/r2b_sweep3_n2.jsonl 2> gpurun_out/r2b_sweep3_n2.err
grep -c '^{' gpurun_out/r2b_sweep3_n*.jsonl
