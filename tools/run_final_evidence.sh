#!/bin/bash
# Round-end evidence on one GPU: full GPU test suite, the bench line, the launch list of the same bench command
# under ncu (gpu__time_duration only), one full ncu capture of the dominant kernel and of the phase kernel.
TAG=${1:-r02}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1; tail -3 gpurun_out/${TAG}_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
tools/ncu_scat1d.sh ${TAG}
cat > /tmp/ncu_phase.py <<'PY'
import sys, os
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import KymatioPhaseScattering1D
from tebscat.synth import ctg_batch
m = KymatioPhaseScattering1D(J=6, Q=8, T=64, shape=4800, device=torch.device('cuda'))
x = ctg_batch(256, 4800, seed=1).cuda()
for _ in range(2):
    m(x, compute_phase=False, compute_cross_phase=True)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:phase_pair_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_phase_tc python /tmp/ncu_phase.py > gpurun_out/ncu_${TAG}_phase.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_phase.log
# the evidence behind the phase kernel's structure: issue rate of tcgen05.mma per issuing thread, MMAs of several
# threads into one accumulator, and the timeline of one CTA (library variant built with -DTEBSCAT_TC_TRACE)
[ -x build/umma_rate ] && ./build/umma_rate > gpurun_out/${TAG}_umma_rate.txt 2>&1
[ -x build/umma_shared_acc ] && ./build/umma_shared_acc > gpurun_out/${TAG}_umma_shared_acc.txt 2>&1
[ -f build/variants/lib_trace.so ] && TEBSCAT_LIB=$PWD/build/variants/lib_trace.so python tools/tc_trace.py > gpurun_out/${TAG}_phase_tc_timeline.txt 2>&1
python tools/time_phase_stages.py 8192 > gpurun_out/${TAG}_phase_stages.txt 2>&1
# BASELINE configs[3] at one GPU, the A/B of the two U0 layouts, the per-step cycles of the headline schedule and the
# per-op breakdown of the large-support level at its two ends
python tools/sweep_config3.py > gpurun_out/${TAG}_sweep_config3_n1.jsonl 2> gpurun_out/${TAG}_sweep_config3_n1.err
AB_SHORT=1 python tools/ab_u0_scratch.py 16384 0,1,2,3 > gpurun_out/${TAG}_ab_u0_layouts.txt 2>&1
python tools/step_profile.py H > gpurun_out/${TAG}_step_cycles_H.json 2> /dev/null
python tools/large_breakdown.py 8 8192 2048 > gpurun_out/${TAG}_breakdown_2_14.txt 2>&1
python tools/large_breakdown.py 10 65536 256 > gpurun_out/${TAG}_breakdown_2_17.txt 2>&1
