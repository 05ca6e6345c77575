for mb in 16 32 64 128 100000; do echo "L2 budget $mb MB"; TEBSCAT_LARGE_L2_MB=$mb SWEEP_POINTS=2,3,5 python tools/sweep_config3.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  J_pad',d['J_pad'],'%.0f sig/s  %.2f TF/s'%(d['signals_per_s'],d['ref_equiv_tflops']))"; done
echo backward; for mb in 16 64 100000; do TEBSCAT_LARGE_L2_MB=$mb python - <<'PY'
import sys,os,time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/vae-teb_b200')
import torch
from tebscat import Scattering1D
S=Scattering1D(6,4800,8,T=64).cuda()
x=torch.randn(2048,4800,device='cuda',requires_grad=True)
for _ in range(3):
    x.grad=None; o,_=S(x); o.backward(torch.ones_like(o))
torch.cuda.synchronize(); t=time.time()
for _ in range(3):
    x.grad=None; o,_=S(x); o.backward(torch.ones_like(o))
torch.cuda.synchronize(); print('  L2 MB',os.environ['TEBSCAT_LARGE_L2_MB'],'fwd+bwd %.0f signals/s'%(3*2048/(time.time()-t)))
PY
done
