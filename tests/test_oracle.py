"""CPU suite: pins the oracle (oracle/kmbart_oracle.py) against
  (1) the committed golden vectors produced by the REFERENCE's own src/model code
      (tests/golden/make_golden.py, via oracle/hf302_shim.py),
  (2) a live run of the reference where /root/reference exists (build container only),
  (3) the installed transformers' BartForConditionalGeneration (independent implementation
      with the same state-dict keys; SURVEY.md §8c "independent cross-checks")."""
import os
import types

import pytest
import torch

from oracle import kmbart_oracle as O
import golden_cases as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmbart_reference_golden.pt")
HAVE_REF = os.path.isdir("/root/reference/src/model")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.fixture(scope="module")
def fwd_case():
    return G.case_forward()


def test_seeded_inputs_reproduce(golden, fwd_case):
    ocfg, sd, batch = fwd_case
    assert torch.allclose(G.checksum(sd, batch), golden["forward"]["checksum"], rtol=1e-12)


def test_forward_loss_logits_match_reference(golden, fwd_case):
    ocfg, sd, batch = fwd_case
    loss, logits, h, enc = O.forward_conditional_generation(sd, ocfg, **batch)
    g = golden["forward"]
    assert abs(loss.item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    assert torch.allclose(enc, g["enc"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(logits[..., G.LOGIT_COLS], g["logits_cols"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(torch.logsumexp(logits, -1), g["logits_lse"], atol=1e-5, rtol=1e-6)
    assert torch.equal(logits.argmax(-1), g["logits_argmax"])


def test_backward_matches_reference(golden, fwd_case):
    ocfg, sd, batch = fwd_case
    osd = {k: v.clone().requires_grad_(k != "final_logits_bias") for k, v in sd.items()}
    loss, _, _, _ = O.forward_conditional_generation(osd, ocfg, **batch)
    loss.backward()
    g = golden["forward"]
    for n, ref in g["grad_norms"].items():
        got = osd[n].grad.norm()
        assert abs(got - ref) <= 1e-4 * ref + 1e-9, n
    for n, ref in g["grad_slices"].items():
        assert torch.allclose(osd[n].grad.reshape(-1)[:64], ref, atol=1e-7, rtol=1e-4), n


def test_visual_rows_give_zero_token_gradient(fwd_case):
    """SURVEY §4 invariant 6: visual rows overwrite token rows, so <img_feat> gets no encoder-input gradient."""
    ocfg, sd, batch = fwd_case
    w = sd["model.shared.weight"].clone().requires_grad_(True)
    osd = dict(sd)
    osd["model.shared.weight"] = w
    enc = O.encoder_forward(osd, ocfg, batch["input_ids"], batch["image_features"], batch["attention_mask"])
    enc.sum().backward()
    assert float(w.grad[ocfg.img_feat_id].abs().max()) == 0.0


def test_cached_default_step_matches_reference(golden, fwd_case):
    """forward() without labels and use_cache=None runs ONE cached step on the last decoder token."""
    ocfg, sd, batch = fwd_case
    enc = O.encoder_forward(sd, ocfg, batch["input_ids"], batch["image_features"], batch["attention_mask"])
    h, caches = O.decoder_forward(sd, ocfg, batch["decoder_input_ids"], enc, batch["attention_mask"], None, None, None, use_cache=True)
    logits = O.lm_logits(sd, h)
    assert logits.shape[1] == 1 and len(caches) == ocfg.decoder_layers
    assert torch.allclose(logits[..., G.LOGIT_COLS], golden["forward"]["cached_default_logits_cols"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name", list(G.GENERATE_CASES))
def test_generate_tokens_match_reference(golden, fwd_case, name):
    ocfg, sd, batch = fwd_case
    torch.manual_seed(G.SAMPLE_SEED)
    toks = O.generate(sd, ocfg, batch["input_ids"], batch["image_features"], batch["attention_mask"], **G.GENERATE_CASES[name])
    assert torch.equal(toks, golden["generate"][name]), name


def test_generate_cache_equals_nocache(golden, fwd_case):
    ocfg, sd, batch = fwd_case
    toks = O.generate(sd, ocfg, batch["input_ids"], batch["image_features"], batch["attention_mask"], use_cache=False,
                      **G.GENERATE_CASES["greedy"])
    assert torch.equal(toks, golden["generate"]["greedy"])
    assert torch.equal(golden["generate"]["greedy_nocache"], golden["generate"]["greedy"])


def test_beam_output_starts_with_forced_bos(golden):
    """SURVEY §4 invariant 8 (src/model/mixins.py:400-402)."""
    for name in ("beam3_early", "beam4_ret2", "beam2_lenpen"):
        t = golden["generate"][name]
        assert (t[:, 0] == 0).all() and (t[:, 1] == 0).all()


def test_pretraining_losses_match_reference(golden):
    pcfg, psd, pbatch = G.case_pretrain()
    assert torch.allclose(G.checksum(psd, pbatch), golden["pretrain"]["checksum"], rtol=1e-12)
    losses, logits = O.forward_pretraining(psd, pcfg, **pbatch)
    for k, ref in golden["pretrain"]["losses"].items():
        assert abs(float(losses[k]) - float(ref)) <= 2e-6 * abs(float(ref)), k
    assert torch.allclose(torch.logsumexp(logits, -1), golden["pretrain"]["logits_lse"], atol=1e-5, rtol=1e-6)


def test_pretraining_requires_mrm_mask():
    pcfg, psd, pbatch = G.case_pretrain()
    pbatch = dict(pbatch, mrm_mask=None)
    with pytest.raises(ValueError):
        O.forward_pretraining(psd, pcfg, **pbatch)


def test_adamw_matches_reference(golden):
    params, grads_seq = G.case_adamw()
    ps = [p.clone() for p in params]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for t, grads in enumerate(grads_seq, 1):
        O.adamw_step(ps, grads, m, v, t, lr=G.ADAMW["lr"], weight_decay=G.ADAMW["weight_decay"])
    for a, b in zip(ps, golden["adamw"]["params"]):
        assert torch.allclose(a, b, atol=1e-7, rtol=1e-6)
    for a, b in zip(v, golden["adamw"]["exp_avg_sq"]):
        assert torch.allclose(a, b, atol=1e-9, rtol=1e-6)


def test_padding_invariance(fwd_case):
    """SURVEY §4 invariant 4: extra pad columns with attention_mask=0 leave non-pad outputs unchanged."""
    ocfg, sd, batch = fwd_case
    _, logits, _, enc = O.forward_conditional_generation(sd, ocfg, **batch)
    B, S = batch["input_ids"].shape
    b2 = dict(batch)
    b2["input_ids"] = torch.cat([batch["input_ids"], torch.full((B, 3), ocfg.pad_token_id)], 1)
    b2["attention_mask"] = torch.cat([batch["attention_mask"], torch.zeros(B, 3, dtype=torch.long)], 1)
    _, logits2, _, enc2 = O.forward_conditional_generation(sd, ocfg, **b2)
    keep = batch["attention_mask"].bool()
    assert torch.allclose(enc2[:, :S][keep], enc[keep], atol=1e-5)
    assert torch.allclose(logits2, logits, atol=2e-5)


def test_random_init_loss_near_log_vocab():
    """SURVEY §4 invariant 2: N(0, 0.02) init gives loss ~ ln(50320) = 10.83."""
    ocfg = G.small_config()
    sd = O.init_state_dict(ocfg, seed=2)
    batch = O.synthetic_batch(ocfg, batch=2, n_regions=4, n_ctx=10, tgt_len=6, seed=8)
    loss, _, _, _ = O.forward_conditional_generation(sd, ocfg, **batch)
    assert abs(loss.item() - 10.826) < 0.3


def test_state_dict_layout_and_param_count():
    """key set / shapes / base parameter count 141 039 360 (+3 791 171 heads)."""
    shapes = O.param_shapes(O.base_config())
    n = sum(int(torch.Size(s).numel()) for s in shapes.values())
    assert n == 141_039_360
    pshapes = O.param_shapes(O.pretrain_base_config(), pretraining=True)
    assert sum(int(torch.Size(s).numel()) for s in pshapes.values()) - n == 3_791_171
    assert shapes["model.encoder.embed_images.linear.weight"] == (768, 2052)
    assert shapes["model.encoder.embed_positions.weight"] == (1026, 768)
    assert "lm_head.weight" not in shapes


# ------------------------------------------------------------------ independent cross-check: transformers 5.x BART
def test_oracle_matches_modern_hf_bart():
    tr = pytest.importorskip("transformers")
    from transformers import BartConfig, BartForConditionalGeneration
    ocfg = G.small_config(vocab_size=50320)
    sd = G.perturb(O.init_state_dict(ocfg, seed=0))
    batch = O.synthetic_batch(ocfg, batch=3, n_regions=5, n_ctx=12, tgt_len=7, seed=12, ragged=True)
    cfg = BartConfig(vocab_size=ocfg.vocab_size, d_model=ocfg.d_model, encoder_layers=2, decoder_layers=2,
                     encoder_attention_heads=2, decoder_attention_heads=2, encoder_ffn_dim=256, decoder_ffn_dim=256,
                     max_position_embeddings=ocfg.max_position_embeddings, dropout=0.0, attention_dropout=0.0,
                     activation_dropout=0.0, activation_function="gelu", scale_embedding=False, pad_token_id=1,
                     bos_token_id=0, eos_token_id=2, decoder_start_token_id=0)
    hf = BartForConditionalGeneration(cfg).eval()
    missing, unexpected = hf.load_state_dict({k: v for k, v in O.full_state_dict(sd).items() if "embed_images" not in k}, strict=False)
    assert not [m for m in missing if "lm_head" not in m], missing
    emb = O.embed_multimodal(sd, ocfg, batch["input_ids"], batch["image_features"])
    with torch.no_grad():
        out = hf(inputs_embeds=emb, attention_mask=batch["attention_mask"], decoder_input_ids=batch["decoder_input_ids"],
                 decoder_attention_mask=batch["decoder_attention_mask"], use_cache=False)
    _, logits, _, enc = O.forward_conditional_generation(sd, ocfg, **batch)
    keep = batch["attention_mask"].bool()
    assert torch.allclose(out.encoder_last_hidden_state[keep], enc[keep], atol=2e-5)
    dkeep = batch["decoder_attention_mask"].bool()
    assert torch.allclose(out.logits[dkeep], logits[dkeep], atol=1e-4)


# ------------------------------------------------------------------ live reference (build container only)
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
def test_live_reference_matches_oracle_on_fresh_seed():
    from oracle import hf302_shim as S
    mods = S.import_reference()
    ocfg = G.small_config(encoder_layers=1, decoder_layers=2)
    sd = G.perturb(O.init_state_dict(ocfg, seed=42), seed=43)
    batch = O.synthetic_batch(ocfg, batch=3, n_regions=5, n_ctx=11, tgt_len=6, seed=44, ragged=True)
    cfg = mods["config"].MultiModalBartConfig(**G.config_kwargs(ocfg))
    model = mods["model"].MultiModalBartForConditionalGeneration(cfg).eval()
    full = O.full_state_dict(sd)
    with torch.no_grad():
        for n, t in model.state_dict().items():
            t.copy_(full[n])
        res = model(**batch)
    loss, logits, _, enc = O.forward_conditional_generation(sd, ocfg, **batch)
    assert abs(res[0].item() - loss.item()) < 1e-6
    assert torch.allclose(res[1], logits, atol=1e-5)
    kw = dict(max_length=6, num_beams=2, early_stopping=True)
    t_ref = model.generate(input_ids=batch["input_ids"], image_features=batch["image_features"],
                           attention_mask=batch["attention_mask"], **kw)
    t_or = O.generate(sd, ocfg, batch["input_ids"], batch["image_features"], batch["attention_mask"], **kw)
    assert torch.equal(t_ref, t_or)
    tok = types.SimpleNamespace(bos_token_id=0, eos_token_id=2, pad_token_id=1)
    assert mods["utils"].sample_sentence is not None and tok.bos_token_id == 0
