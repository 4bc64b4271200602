"""Dev probe (GPU): torch.profiler breakdown of one data-parallel training step (config 3: 256x512 crops, batch 8)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "explicit-context-mapping-for-stereo-matching_b200"))
from cmf.models import get_model  # noqa: E402
from cmf_b200 import parallel as par  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(3)
left, right = torch.rand(nb, 3, 256, 512, generator=g).to(dev), torch.rand(nb, 3, 256, 512, generator=g).to(dev)
disp = (torch.rand(nb, 256, 512, generator=g) * 230 - 10).to(dev)
torch.manual_seed(0)
model = get_model("cmfsm").to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
for _ in range(2):
    par.dp_train_step(model, opt, left, right, disp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
par.dp_train_step(model, opt, left, right, disp)
e1.record()
torch.cuda.synchronize()
print("step %.1f ms" % e0.elapsed_time(e1))
# forward only
with torch.no_grad():
    e0.record()
    model(left, right)
    e1.record()
    torch.cuda.synchronize()
print("forward (no_grad, train mode) %.1f ms" % e0.elapsed_time(e1))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    par.dp_train_step(model, opt, left, right, disp)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
