import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from fqss_b200 import tcn_engine as E
DEV = "cuda"
torch.manual_seed(3)
B, Ci, Co, M = 3, 128, 256, 1003
qmin, qmax = torch.tensor([-1.3], device=DEV), torch.tensor([2.1], device=DEV)
delta = (qmax - qmin) / 255
codes = torch.randint(0, 256, (B, Ci, M), device=DEV).float()
x = (delta * codes + qmin)
W = (torch.randn(Co, Ci, 1, device=DEV) * 0.1)
bias = (torch.randn(Co, device=DEV) * 0.1)
wmax = W.amax(dim=(1, 2), keepdim=True).clone()
wmin = W.amin(dim=(1, 2), keepdim=True).clone()
y = E.CodeConv1x1.apply(x, qmin, qmax, W, wmin, wmax, bias)
a = torch.maximum(wmin.abs(), wmax.abs()); dl = 2 * a / 255; t = W / dl
Wq = dl * torch.clamp(torch.round(t), -128, 127)
yr = F.conv1d(x.double(), Wq.double(), bias.double()).float()
d = (y - yr).abs()
print("max abs", d.max().item(), "mean abs", d.mean().item(), "yr absmax", yr.abs().max().item())
idx = (d > 1e-3).nonzero()
print("n bad", idx.shape[0], "of", d.numel())
print("bad by sample", [(idx[:, 0] == b).sum().item() for b in range(B)])
print("bad channels (first 10 unique)", idx[:, 1].unique()[:10].tolist(), "n unique ch", idx[:, 1].unique().numel())
print("bad frames unique n", idx[:, 2].unique().numel(), idx[:, 2].unique()[:10].tolist(), idx[:, 2].unique()[-10:].tolist())
# without bias / zero-point: check pieces
y2 = E.CodeConv1x1.apply(x, qmin, qmax, W, wmin, wmax, None)
print("no-bias diff", (y2 - (yr - bias.view(1, -1, 1))).abs().max().item())
from fqss_b200._native import lib, check, ptr, stream_ptr
L = E._libx()
bf = torch.bfloat16
Wc, WcT = torch.empty((Co, Ci), dtype=bf, device=DEV), torch.empty((Ci, Co), dtype=bf, device=DEV)
s1, s0, dws = torch.empty(Co, device=DEV), torch.empty(Co, device=DEV), torch.empty(Co, device=DEV)
Wf = W.reshape(Co, Ci)
check(L.fqss_tcn_prep(ptr(Wf), ptr(wmin), ptr(wmax), None, ptr(qmin), ptr(qmax), ptr(Wc), ptr(WcT), ptr(s1), ptr(s0), ptr(dws), Co, Ci, Co, 0, 0, stream_ptr()))
torch.cuda.synchronize()
cref = torch.clamp(torch.round(t), -128, 127).reshape(Co, Ci)
dc = (Wc.float() - cref)
print("weight code mismatches", (dc != 0).sum().item(), "channels", (dc != 0).any(1).sum().item())
bad = (dc != 0).nonzero()[:5]
for o, i in bad.tolist():
    print(o, i, "prep", Wc[o, i].item(), "ref", cref[o, i].item(), "t", t.reshape(Co, Ci)[o, i].item(), "W", Wf[o, i].item(), "dl", dl.reshape(-1)[o].item(), "dws", dws[o].item())
print("dws vs dl", (dws - dl.reshape(-1)).abs().max().item())
