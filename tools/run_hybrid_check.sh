#!/bin/bash
# Large-support level after the fused subtrees: GPU tests of the level, BASELINE configs[3] points 2..5 with and
# without the fused subtrees, per-op breakdown at the two ends.
TAG=${1:-r02h}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_large.py -x -q > gpurun_out/${TAG}_gpu_large.log 2>&1; tail -5 gpurun_out/${TAG}_gpu_large.log
python -m pytest tests/test_gpu_phase.py -x -q -k "2_14" > gpurun_out/${TAG}_gpu_phase_large.log 2>&1; tail -2 gpurun_out/${TAG}_gpu_phase_large.log
SWEEP_POINTS=2,3,4,5 python tools/sweep_config3.py > gpurun_out/${TAG}_sweep_hybrid.jsonl 2> gpurun_out/${TAG}_sweep_hybrid.err
SWEEP_POINTS=2,3,4,5 TEBSCAT_HYBRID=0 python tools/sweep_config3.py > gpurun_out/${TAG}_sweep_nohybrid.jsonl 2> gpurun_out/${TAG}_sweep_nohybrid.err
python tools/large_breakdown.py 8 8192 2048 > gpurun_out/${TAG}_breakdown_2_14.txt 2>&1
python tools/large_breakdown.py 10 65536 256 > gpurun_out/${TAG}_breakdown_2_17.txt 2>&1
cat gpurun_out/${TAG}_sweep_hybrid.jsonl gpurun_out/${TAG}_sweep_nohybrid.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['J'], d['N'], d.get('level'), round(d['signals_per_s']), round(d['ref_equiv_tflops'], 2))"
head -12 gpurun_out/${TAG}_breakdown_2_14.txt; head -14 gpurun_out/${TAG}_breakdown_2_17.txt
