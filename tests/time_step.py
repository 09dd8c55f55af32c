"""Quick A/B timer: config-2 train step (device-resident batch), CUDA events.  python tests/time_step.py [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
import bench
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.optim import AdamW

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
large = len(sys.argv) > 2 and sys.argv[2] == "large"   # BASELINE configs[4]: d=1024 12+12, 100 RoIs + 256 ctx tokens, batch 64/GPU
if large:
    from kmbart.synth import synthetic_batch
    cfg = MultiModalBartConfig(max_position_embeddings=1024)
    BATCH, FLOP_PER_SAMPLE = 64, 464.662e9
else:
    cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
    BATCH, FLOP_PER_SAMPLE = 128, 56.412e9
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().train()
opt = AdamW(model.parameters(), lr=1e-5)
if large:
    b = synthetic_batch(cfg, batch=BATCH, n_regions=100, n_ctx=256, tgt_len=48, seed=1234)
    batch = {k: ([t.cuda() for t in v] if isinstance(v, list) else v.cuda()) for k, v in b.items()}
else:
    batch = bench.make_batch(cfg, 1234, device="cuda")

def step():
    loss = model(**batch)[0]
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss

for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train step {ms:.3f} ms  ({BATCH / ms * 1e3:.0f} samples/s, {BATCH * FLOP_PER_SAMPLE / ms / 1e9:.0f} TFLOP/s)  loss {loss.item():.5f}  "
      f"PDL={'off' if os.environ.get('KMBART_NO_PDL') == '1' else 'on'}  config={'large d=1024 12+12 S_e=356 B=64' if large else 'base B=128'}")
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host issue time {1e3 * (t1 - t0) / steps:.3f} ms/step")
# host-side issue cost of ONE step with an empty queue (no back-pressure from the GPU)
acc = [0.0, 0.0, 0.0]
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); loss = model(**batch)[0]; t1 = time.perf_counter()
    opt.zero_grad(); loss.backward(); t2 = time.perf_counter()
    opt.step(); t3 = time.perf_counter()
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
print("host issue (empty queue) ms: forward %.3f  backward %.3f  optimizer %.3f" % tuple(1e3 * a / 5 for a in acc))
