"""Launch timeline of one data-parallel training step (configs[1] shapes) on rank 0: every kernel and memcpy with its
stream, start and duration (torch.profiler / CUPTI), written as CSV.  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/prof_timeline_dp.py out.csv"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
import torch.distributed as dist
import bench
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.optim import AdamW
from kmbart.parallel import FlatGradReducer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_dp.csv"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().train()
opt = AdamW(model.parameters(), lr=1e-5)
if world > 1:
    red = FlatGradReducer(model, defer_tail=True)
batch = bench.make_batch(cfg, 1234 + rank, device="cuda")

def step():
    loss = model(**batch)[0]
    opt.zero_grad()
    loss.backward()
    opt.step()

for _ in range(6):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in evs:
            f.write(f"{e.time_range.start - t0:.1f},{e.time_range.end - e.time_range.start:.1f},{getattr(e, 'stream', -1) if hasattr(e, 'stream') else -1},\"{e.name[:90]}\"\n")
    print("events", len(evs), "transport", getattr(red, "transport", None) if world > 1 else None)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
