#!/bin/bash
# ncu capture of the fused cascade kernel on a 4096-signal batch: full set with source (for --page source) and a raw
# metrics CSV.  Usage (on the GPU box): tools/ncu_scat1d.sh <tag>
set -e
TAG=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/ncu_drv.py <<'PY'
import sys, os
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import Scattering1D
from tebscat.synth import ctg_batch
S = Scattering1D(6, 4800, 8, T=64).cuda()
x = ctg_batch(256, 4800, seed=1).repeat(8, 1, 1).reshape(-1, 4800).cuda().contiguous()
for _ in range(3):
    out, _ = S(x)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:scat1d_kernel -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_scat1d python /tmp/ncu_drv.py > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_scat1d.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_scat1d_raw.csv 2>/dev/null || true
