import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fqss_oracle as O
from fqss_b200.testing import FULL_CFG, FULL_KW, model_pair, oracle_params
from fqss_b200.qat.models.convtasnetq import ConvTasNetQ
from fqss_b200 import ops
DEV = "cuda"
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
model, fmodel = model_pair(FULL_KW, DEV, seed=0)
gen = torch.Generator().manual_seed(1)
src = torch.randn(1, 2, 32000, generator=gen) * 0.05
mix = src.sum(1, keepdim=True)
fP = oracle_params(fmodel)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    fo = O.separator_forward(fP, mix, FULL_CFG, quant=False)
    ConvTasNetQ.use_float_engine = False
    ft = fmodel(mix.to(DEV))
    ConvTasNetQ.use_float_engine = True
    fe = fmodel(mix.to(DEV))
    torch.backends.cudnn.allow_tf32 = True
    ConvTasNetQ.use_float_engine = False
    ftf = fmodel(mix.to(DEV))
    ConvTasNetQ.use_float_engine = True
print("teacher: engine vs oracle %.3e | torch fp32 vs oracle %.3e | torch tf32 vs oracle %.3e | engine vs torch %.3e" % (rel(fe, fo), rel(ft, fo), rel(ftf, fo), rel(fe, ft)))
print("norms", fo.norm().item(), fe.norm().item())
est = torch.randn(1, 2, 32000, generator=gen) * 0.05
for name, f in (("oracle", fo), ("engine", fe.cpu()), ("torch", ft.cpu()), ("tf32", ftf.cpu())):
    lo, _ = O.fqss_kd_loss(est, f, src, 0.1)
    lc = ops.kd_loss(est.to(DEV), f.to(DEV), src.to(DEV), 0.1)
    print(name, "loss oracle-arith %.5f  cuda-kernel %.5f" % (lo.item(), lc[0].item()))
