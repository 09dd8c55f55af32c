#!/usr/bin/env python
"""Generates tests/golden/*.pt by running the REFERENCE's own src/model code (imported from
/root/reference through oracle/hf302_shim.py, which stands in for the un-installable
transformers==3.0.2) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures hold outputs (and checksums of the seeded inputs/weights), not the inputs: every
consumer regenerates weights and batches from the recorded seeds with
oracle.kmbart_oracle.init_state_dict / synthetic_batch and compares its result with what the
reference produced.  tests/test_oracle.py checks the CPU oracle against them; the `-m gpu`
tests check the CUDA path against them.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import hf302_shim as S  # noqa: E402
from oracle import kmbart_oracle as O  # noqa: E402
import golden_cases as G  # noqa: E402


def build_reference(mods, ocfg, sd, pretraining=False):
    cfg = mods["config"].MultiModalBartConfig(**G.config_kwargs(ocfg))
    cls = mods["model"].MultiModalBartForPreTraining if pretraining else mods["model"].MultiModalBartForConditionalGeneration
    model = cls(cfg).eval()
    full = O.full_state_dict(sd)
    with torch.no_grad():
        for n, t in model.state_dict().items():   # reference load_state_dict() is unusable (custom _load_from_state_dict)
            t.copy_(full[n])
    assert [n for n, _ in model.named_parameters()] == list(O.param_shapes(ocfg, pretraining)), "parameters() order"
    return model


def main():
    mods = S.import_reference()
    out = {}

    # ---- 1. fine-tuning forward + loss + gradients (src/model/model.py:325-405)
    ocfg, sd, batch = G.case_forward()
    model = build_reference(mods, ocfg, sd)
    for p in model.parameters():
        p.requires_grad_(True)
    res = model(**batch)
    loss, logits, enc = res[0], res[1], res[-1]
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    out["forward"] = dict(
        checksum=G.checksum(sd, batch), loss=loss.detach(), enc=enc.detach(),
        logits_cols=logits.detach()[..., G.LOGIT_COLS], logits_lse=torch.logsumexp(logits.detach(), -1),
        logits_argmax=logits.detach().argmax(-1),
        grad_norms={n: g.norm() for n, g in grads.items()},
        grad_slices={n: grads[n].reshape(-1)[:64].clone() for n in G.GRAD_SLICE_NAMES},
    )
    # logits without labels are outputs[0] (scripts/filter_reason.py:42)
    with torch.no_grad():
        nb = {k: v for k, v in batch.items() if k != "labels"}
        res2 = model(use_cache=False, **nb)
        assert torch.equal(res2[0], logits.detach())
        # use_cache=None -> config.use_cache=True: ONE cached step on the last decoder token (src/model/model.py:57-70)
        res3 = model(**nb)
    assert res3[0].shape[1] == 1 and len(res3) == 3
    out["forward"]["cached_default_logits_cols"] = res3[0][..., G.LOGIT_COLS].clone()

    # ---- 2. generation (src/model/mixins.py:33-384 + inherited loops)
    model.zero_grad()
    gen = {}
    gb = dict(input_ids=batch["input_ids"], image_features=batch["image_features"], attention_mask=batch["attention_mask"])
    for name, kw in G.GENERATE_CASES.items():
        torch.manual_seed(G.SAMPLE_SEED)
        gen[name] = model.generate(**gb, **kw)
    torch.manual_seed(G.SAMPLE_SEED)
    gen["greedy_nocache"] = model.generate(**gb, use_cache=False, **G.GENERATE_CASES["greedy"])
    assert torch.equal(gen["greedy_nocache"], gen["greedy"]), "cache vs no-cache greedy must agree"
    out["generate"] = gen
    tok = types.SimpleNamespace(bos_token_id=ocfg.bos_token_id, eos_token_id=ocfg.eos_token_id, pad_token_id=ocfg.pad_token_id)
    torch.manual_seed(G.SAMPLE_SEED)
    ids, lp = mods["utils"].sample_sentence(model, batch["input_ids"], batch["image_features"], batch["attention_mask"], tok,
                                            top_k=20, top_p=0.9, max_length=7)
    out["sample_sentence"] = dict(ids=ids, logprobs=lp)

    # ---- 3. multitask pretraining forward (src/model/model.py:162-309)
    pcfg, psd, pbatch = G.case_pretrain()
    pmodel = build_reference(mods, pcfg, psd, pretraining=True)
    for p in pmodel.parameters():
        p.requires_grad_(True)
    pres = pmodel(**pbatch)
    losses = pres[0]
    losses["loss"].backward()
    pg = {n: p.grad for n, p in pmodel.named_parameters() if p.grad is not None}
    out["pretrain"] = dict(checksum=G.checksum(psd, pbatch), losses={k: v.detach() for k, v in losses.items()},
                           logits_lse=torch.logsumexp(pres[1].detach(), -1),
                           grad_norms={n: g.norm() for n, g in pg.items()})

    # ---- 4. AdamW (transformers.AdamW, vcg_train.py:100)
    params, grads_seq = G.case_adamw()
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = S.AdamW(ps, lr=G.ADAMW["lr"], weight_decay=G.ADAMW["weight_decay"])
    for grads_t in grads_seq:
        for p, g in zip(ps, grads_t):
            p.grad = g.clone()
        opt.step()
    out["adamw"] = dict(params=[p.detach().clone() for p in ps],
                        exp_avg=[opt.state[p]["exp_avg"].clone() for p in ps],
                        exp_avg_sq=[opt.state[p]["exp_avg_sq"].clone() for p in ps])

    # ---- 5. partial_load slice copy through the reference's from_pretrained (src/model/mixins.py:511-528)
    import json
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        small_v = 50265
        ck = {k: v.clone() for k, v in O.full_state_dict(sd).items()}
        for k in ("model.shared.weight", "model.encoder.embed_tokens.weight", "model.decoder.embed_tokens.weight"):
            ck[k] = ck[k][:small_v] + 1.0
        ck["final_logits_bias"] = ck["final_logits_bias"][:, :small_v] + 1.0
        torch.save(ck, os.path.join(td, "pytorch_model.bin"))
        cfgd = G.config_kwargs(ocfg)
        cfgd["partial_load"] = ["final_logits_bias", "model.shared.weight", "model.encoder.embed_tokens.weight",
                                "model.decoder.embed_tokens.weight"]
        cfg = mods["config"].MultiModalBartConfig(**cfgd)
        torch.manual_seed(11)
        loaded = mods["model"].MultiModalBartForConditionalGeneration.from_pretrained(td, config=cfg, error_on_mismatch=False)
        w = loaded.model.shared.weight.detach()
        out["partial_load"] = dict(head_equal=bool(torch.equal(w[:small_v], ck["model.shared.weight"])),
                                   tail_untouched=bool((w[small_v:] - 1.0).abs().max() > 0.5),
                                   flb_head=bool(torch.equal(loaded.final_logits_bias[:, :small_v], ck["final_logits_bias"])),
                                   flb_tail_zero=bool((loaded.final_logits_bias[:, small_v:] == 0).all()))

    path = os.path.join(HERE, "kmbart_reference_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out["partial_load"].items():
        print("partial_load", k, v)
    print("losses", {k: float(v) for k, v in out["pretrain"]["losses"].items()}, "ft loss", float(out["forward"]["loss"]))
    for k, v in gen.items():
        print(k, tuple(v.shape))


if __name__ == "__main__":
    main()
