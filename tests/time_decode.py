"""Generation-only timer: bench.bench_generation() (tokens/s for greedy / top-k / beam-5 and the decode-step graph replay
against the HBM roofline) without the training legs.  python tests/time_decode.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
import bench
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration

cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
dev = torch.device("cuda:0")
model = MultiModalBartForConditionalGeneration(cfg).to(dev).train()
burst, sustained, hbm, src = bench.load_peaks()
gen = bench.bench_generation(model, cfg, dev, 0, 1, None, hbm, src)
print(json.dumps({"tokens_per_s": gen["tokens_per_s"], "decode_step": gen["decode_step"]}))
