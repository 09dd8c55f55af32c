"""Profiling driver: config-2 train step; warm-up outside, N profiled steps inside cudaProfilerStart/Stop.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tests/prof_step.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
import bench
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.optim import AdamW

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
large = len(sys.argv) > 2 and sys.argv[2] == "large"   # BASELINE configs[4] shapes, batch 64
if large:
    cfg = MultiModalBartConfig(max_position_embeddings=1024)
else:
    cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().train()
opt = AdamW(model.parameters(), lr=1e-5)
if large:
    from kmbart.synth import synthetic_batch
    b = synthetic_batch(cfg, batch=64, n_regions=100, n_ctx=256, tgt_len=48, seed=1234)
    batch = {k: ([t.cuda() for t in v] if isinstance(v, list) else v.cuda()) for k, v in b.items()}
else:
    batch = bench.make_batch(cfg, 1234, device="cuda")

def step():
    loss = model(**batch)[0]
    opt.zero_grad()
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
print("opt.launches_last", opt.launches_last, "eng.launches_last", model._engine().launches_last)
print("n params with grad", sum(p.grad is not None for p in model.parameters()))
st = opt.state[next(iter(model.parameters()))]
print("state step", st.get("step"), "exp_avg norm", st["exp_avg"].norm().item() if "exp_avg" in st else None)
