"""Which ATen kernels still run inside a steady-state QAT step, and who calls them (torch.profiler with stacks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from fqss_b200.losses import fqss_kd_loss
from fqss_b200.parallel import ParamArena
from fqss_b200.qat.models.load_model import enable_observer
from fqss_b200.testing import FULL_KW, model_pair
dev = torch.device("cuda", 0)
B, T = int(os.environ.get("B", "32")), 32000
model, fmodel = model_pair(FULL_KW, dev, seed=0)
src = torch.randn(B, 2, T, device=dev) * 0.05
mix = src.sum(1, keepdim=True)
with torch.no_grad():
    model(mix[:4]); model(mix[:4])
enable_observer(model, False)
arena = ParamArena(list(model.parameters()))
def step():
    arena.zero_grad()
    est = model(mix)
    with torch.no_grad():
        fest = fmodel(mix)
    loss, _, _ = fqss_kd_loss(est, fest, src, 0.1)
    loss.backward()
    arena.gather_grads()
    s = arena.allreduce_mean()
    arena.clip_and_step(pre_scale=s, max_norm=5.0, lr=1e-3)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 20 and e.name.startswith("aten::")]
evs.sort(key=lambda e: -e.device_time_total)
tot = 0
for e in evs[:14]:
    st = [f for f in (e.stack or []) if "fqss_b200" in f or "bench" in f or "scratch" in f][:4]
    if not st:
        st = list(e.stack or [])[:4]
    print("%8.1f us %-28s %s | %s" % (e.device_time_total, e.name, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in st)))
for e in prof.events():
    if e.name.startswith("aten::") and e.device_time_total > 0 and not e.cpu_children:
        tot += e.device_time_total
print("leaf aten device time total: %.1f us" % tot)
