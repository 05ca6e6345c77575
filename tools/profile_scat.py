"""Small driver for ncu: a few forward passes of the headline config on a reduced batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'vae-teb_b200'))
import torch                                    # noqa: E402
from tebscat import Scattering1D                # noqa: E402
from tebscat.synth import ctg_batch             # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = Scattering1D(6, 4800, 8, T=64).cuda()
x = ctg_batch(max(B // 2, 1), 4800, seed=3).reshape(-1, 4800)[:B].cuda()
for _ in range(reps):
    out, _ = S(x)
torch.cuda.synchronize()
print('ok', tuple(out.shape), float(out.abs().mean()))
