#!/bin/bash
# One full ncu capture of the tcgen05 phase kernel (stage B, dense form).  usage: tools/ncu_phase_tc.sh TAG
TAG=${1:-x}
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
