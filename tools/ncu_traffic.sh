#!/bin/bash
# DRAM / L2 traffic of ONE full-batch launch (16 384 signals) of the headline kernel: the `roofline.traffic` figure
# of bench.py.  usage (on the GPU box): tools/ncu_traffic.sh TAG  ->  gpurun_out/TAG_traffic.csv
TAG=${1:-r02}
cat > /tmp/ncu_traffic.py <<'PY'
import sys, os
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch
from tebscat import Scattering1D
from tebscat.synth import ctg_batch
S = Scattering1D(6, 4800, 8, T=64).cuda()
x = ctg_batch(8192, 4800, seed=1234).reshape(-1, 4800).cuda().contiguous()
for _ in range(3):
    out, _ = S(x)
torch.cuda.synchronize()
PY
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scat1d_kernel -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_traffic.csv python /tmp/ncu_traffic.py > gpurun_out/${TAG}_traffic.log 2>&1
cat gpurun_out/${TAG}_traffic.csv | tail -5
