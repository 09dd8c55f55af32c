"""Profiling driver for the decode chain: greedy generation at the config-4 per-GPU shape (batch 64, 24 new tokens);
the graphs are captured outside the profiled region, the profiled region replays ONE full generate() call.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tests/prof_decode.py [beam]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from kmbart.synth import synthetic_batch, to_device
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration

beam = len(sys.argv) > 1 and sys.argv[1] == "beam"
cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().eval()
b = to_device(synthetic_batch(cfg, batch=64, seed=4321), "cuda")
gi = dict(input_ids=b["input_ids"], image_features=b["image_features"], attention_mask=b["attention_mask"])
kw = dict(max_length=25, min_length=24, num_beams=5, early_stopping=True) if beam else dict(max_length=25, min_length=25)
with torch.no_grad():
    model.generate(**gi, **kw)
    model.generate(**gi, **kw)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    t = model.generate(**gi, **kw)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done", tuple(t.shape))
