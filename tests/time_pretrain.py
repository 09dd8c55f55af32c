"""BASELINE configs[2]: multitask pre-training step (MultiModalBartForPreTraining, config/pretrain_base.json: LM x5 + masked
region KL + attribute CE + relation CE; src/model/model.py:162-309) at batch 128, S_e = 100, S_d = 38 + 48 = 86.
CUDA events, device-resident batch.   python tests/time_pretrain.py [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from kmbart.synth import synthetic_batch
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForPreTraining
from kmbart.optim import AdamW

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "pretrain_base.json"))))
B, R, T = 128, 36, 48
torch.manual_seed(0)
model = MultiModalBartForPreTraining(cfg).cuda().train()
opt = AdamW(model.parameters(), lr=1e-5)
batch = synthetic_batch(cfg, batch=B, n_regions=R, n_ctx=64, tgt_len=T, seed=1234)
g = torch.Generator().manual_seed(9)
Sd = R + 2 + T
dec = torch.full((B, Sd), cfg.pad_token_id, dtype=torch.long)
labels = torch.full((B, Sd), -100, dtype=torch.long)
mrm_mask = torch.zeros(B, Sd, dtype=torch.bool)
attr_mask = torch.zeros(B, Sd, dtype=torch.bool)
mrm_labels, attr_labels, rel_labels = [], [], []
for b in range(B):
    slots = [cfg.img_feat_id] * R
    masked = [i for i in range(R) if torch.rand(1, generator=g).item() < 0.2]          # MRM p = 0.2 (src/data/collation.py:113-132)
    for i in masked:
        slots[i] = cfg.cls_token_id
        mrm_mask[b, 1 + i] = True
        batch["image_features"][b][i, :2048] = 0                                          # feature zeroed, box kept
    dec[b] = torch.tensor([50265] + slots + [50266, cfg.bos_token_id] + batch["decoder_input_ids"][b, 1:T].tolist())
    labels[b, :R + 2] = torch.tensor([cfg.cls_token_id] * (R + 2))
    labels[b, R + 2:] = batch["labels"][b, :T]
    mrm_labels.append(torch.softmax(torch.randn(len(masked), cfg.num_labels, generator=g), -1))
    attr_slots = [i for i in range(R) if i not in masked][:16]
    for i in attr_slots:
        attr_mask[b, 1 + i] = True
    attr_labels.append(torch.randint(0, cfg.num_attributes, (len(attr_slots),), generator=g))
    rels = []
    for _ in range(32):
        o, s = torch.randint(0, R, (2,), generator=g).tolist()
        rels.append({"object_index": 1 + o, "subject_index": 1 + s, "label": int(torch.randint(0, cfg.num_relations, (1,), generator=g))})
    rel_labels.append(rels)
batch.update(decoder_input_ids=dec, labels=labels, decoder_attention_mask=torch.ones(B, Sd, dtype=torch.long),
             mrm_labels=mrm_labels, mrm_mask=mrm_mask, attribute_labels=attr_labels, attribute_mask=attr_mask,
             relation_labels=rel_labels)
dev = {}
for k, v in batch.items():
    if k == "relation_labels":
        dev[k] = v                                   # host dicts, like the reference (src/training.py:45)
    elif isinstance(v, list):
        dev[k] = [t.cuda() for t in v]
    else:
        dev[k] = v.cuda()

def step():
    out = model(**dev)
    loss = out[0]["loss"]
    opt.zero_grad()
    loss.backward()
    opt.step()
    return out[0]

for _ in range(4):
    l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
FLOP = 77.012e9   # per sample per step, SURVEY.md §8d config 3
print(f"pretrain step {ms:.3f} ms  ({B / ms * 1e3:.0f} samples/s, {B * FLOP / ms / 1e9:.0f} TFLOP/s)  " +
      "  ".join(f"{k}={float(v):.4f}" for k, v in l.items()))
