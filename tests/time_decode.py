"""Decode-step timing at the configs[3] per-GPU shape (batch 64, S_e = 100, 24 new tokens): CUDA-graph replays of the
whole step (model chain + LM head [+ selection]) for every implementation of the step —
  mega    csrc/decode_mega.cu      one persistent launch, 11 grid-barrier phases per layer (default)
  cluster csrc/decode_cluster.cu   one persistent launch of 4-CTA clusters, 6 phases per layer (KMBART_DECODE_CLUSTER=1)
  chain   the 68-kernel launch chain with tcgen05 GEMMs (KMBART_DECODE_CHAIN=1)
at rows 64 (greedy, device-side selection in the graph) and rows 320 (beam 5: model chain + LM head only).
    python tests/time_decode.py [impl ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.decode import DecodeSession

impls = sys.argv[1:] or ["mega", "cluster", "chain"]
cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().eval()
eng = model._engine()
eng.sync_shadow()
B, Se, NEW = 64, 100, 24
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for impl in impls:
    os.environ["KMBART_DECODE_CHAIN"] = "1" if impl == "chain" else "0"
    os.environ["KMBART_DECODE_CLUSTER"] = "1" if impl == "cluster" else "0"
    for label, rows, sel, use_tbl in (("rows64_greedy", B, dict(do_sample=False, temperature=1.0, top_k=50, eos=cfg.eos_token_id, pad=cfg.pad_token_id,
                                                                  min_length=NEW + 1), False), ("rows320_beam5", B * 5, None, True)):
        sess = DecodeSession(eng, B, Se, rows, NEW + 1, False)
        sess.begin(torch.randn(B, Se, cfg.d_model, device="cuda") * 0.5, None, cfg.decoder_start_token_id, use_tbl)

        def run():
            for t in range(NEW):
                sess.step(t, model.final_logits_bias, sel)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * NEW)
        nbytes = sum(sess.step_bytes(t) for t in range(NEW)) / NEW
        print(f"{impl:8s} {label:14s} {us:8.1f} us/step  {nbytes / us / 1e3:7.1f} GB/s  frac_of_hbm_peak {nbytes / us / 1e3 / hbm:.3f}  kernels/step {sess.launches_per_step}")
        del sess
