"""CPU oracle for the KM-BART hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference`
legs may import this module.  The product path (km-bart_b200/) never does; it fails
loudly when libkmbart_sm100.so is missing.

What this is: a functional, fp32, pure-PyTorch restatement of the arithmetic that
fomalhautb/KM-BART executes for forward / loss / generate, written against a flat
``state_dict`` (same key names as the reference) instead of
nn.Modules.  The transformer arithmetic itself lives in the reference's un-vendored
dependency ``transformers==3.0.2`` (environment.yaml:159; modules modeling_bart,
generation_utils, optimization) which is absent from /root/reference and cannot be
installed offline, so its published algorithm is restated here and anchored on the
reference's own call sites:

  encoder embed / merge ........ src/model/modules.py:24-41, :89-102, :133-137
  encoder stack ................ src/model/modules.py:140-165 (HF-3.0.2 EncoderLayer)
  decoder inputs / masks ....... src/model/model.py:63-70   (HF-3.0.2 _prepare_bart_decoder_inputs)
  decoder stack + KV cache ..... src/model/model.py:87-97   (HF-3.0.2 BartDecoder/DecoderLayer/SelfAttention)
  LM head + CE (fine-tune) ..... src/model/model.py:397-403
  LM head + multitask losses ... src/model/model.py:244-309
  generate front half .......... src/model/mixins.py:33-384
  BOS/EOS forcing .............. src/model/mixins.py:400-417
  beam cache re-order .......... src/model/mixins.py:419-434
  greedy/sample/beam loops ..... HF-3.0.2 generation_utils (reached from mixins.py:336-382)
  AdamW ........................ HF-3.0.2 optimization.AdamW (constructed vcg_train.py:100)

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path
 and cannot be imported unmodified in this container.  The oracle is
pinned two ways instead (tests/test_oracle.py):
  1. against the reference's *own* src/model code, imported from /root/reference with
     oracle/hf302_shim.py standing in for the missing transformers-3.0.2 symbols —
     fixtures in tests/golden/ are produced by tests/golden/make_golden.py that way;
  2. against the installed transformers-5.5 BartForConditionalGeneration (same
     state-dict keys, independent implementation of the same math).
Because (1) still relies on a restated dependency, DESIGN.md reports parity as
"pinned to reference src/model + restated HF-3.0.2", not to an unmodified reference run.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# ----------------------------------------------------------------------------- config
class OracleConfig:
    """The BartConfig / MultiModalBartConfig fields the path reads (src/model/config.py:4-92)."""

    def __init__(self, **kw):
        d = dict(
            vocab_size=50320, d_model=1024, image_feature_size=2052,
            encoder_layers=12, decoder_layers=12,
            encoder_attention_heads=16, decoder_attention_heads=16,
            encoder_ffn_dim=4096, decoder_ffn_dim=4096,
            max_position_embeddings=1024, extra_pos_embeddings=2,
            dropout=0.1, attention_dropout=0.0, activation_dropout=0.0, classif_dropout=0.0,
            init_std=0.02, pad_token_id=1, bos_token_id=0, eos_token_id=2,
            decoder_start_token_id=0, img_feat_id=50273, cls_token_id=50276,
            scale_embedding=False, normalize_embedding=True, normalize_before=False,
            add_final_layer_norm=False, num_labels=1, num_attributes=1, num_relations=1,
            lm_loss_factor=1.0, mrm_loss_factor=1.0, attribute_loss_factor=1.0,
            relation_loss_factor=1.0,
            # PretrainedConfig generation defaults (HF-3.0.2 configuration_utils)
            max_length=20, min_length=0, do_sample=False, early_stopping=False, num_beams=1,
            temperature=1.0, top_k=50, top_p=1.0, repetition_penalty=1.0, length_penalty=1.0,
            no_repeat_ngram_size=0, num_return_sequences=1, use_cache=True,
        )
        d.update(kw)
        self.__dict__.update(d)

    @classmethod
    def from_json(cls, path):
        import json
        with open(path) as f:
            return cls(**json.load(f))


def base_config(**kw):
    """config/vcg_base.json of the reference (bart-base 6+6, d=768)."""
    d = dict(d_model=768, encoder_layers=6, decoder_layers=6, encoder_attention_heads=12,
             decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072)
    d.update(kw)
    return OracleConfig(**d)


def pretrain_base_config(**kw):
    """config/pretrain_base.json of the reference."""
    d = dict(num_labels=1601, num_attributes=129, num_relations=129, lm_loss_factor=5,
             mrm_loss_factor=1, attribute_loss_factor=1, relation_loss_factor=1)
    d.update(kw)
    return base_config(**d)


# ----------------------------------------------------------------------------- parameters
def param_shapes(cfg, pretraining=False) -> Dict[str, Tuple[int, ...]]:
    """Unique (untied) parameter tensors in `parameters()` order of the reference modules.

    model.encoder.embed_tokens.weight and model.decoder.embed_tokens.weight alias
    model.shared.weight (src/model/model.py:32-35)."""
    d, V = cfg.d_model, cfg.vocab_size
    npos = cfg.max_position_embeddings + cfg.extra_pos_embeddings
    s = {}
    s["model.shared.weight"] = (V, d)
    s["model.encoder.embed_images.linear.weight"] = (d, cfg.image_feature_size)
    s["model.encoder.embed_images.linear.bias"] = (d,)
    s["model.encoder.embed_positions.weight"] = (npos, d)

    def attn(p):
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[f"{p}.{n}.weight"] = (d, d)
            s[f"{p}.{n}.bias"] = (d,)

    def ln(p):
        s[f"{p}.weight"] = (d,)
        s[f"{p}.bias"] = (d,)

    for i in range(cfg.encoder_layers):
        p = f"model.encoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        s[p + ".fc1.weight"] = (cfg.encoder_ffn_dim, d)
        s[p + ".fc1.bias"] = (cfg.encoder_ffn_dim,)
        s[p + ".fc2.weight"] = (d, cfg.encoder_ffn_dim)
        s[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")
    ln("model.encoder.layernorm_embedding")
    s["model.decoder.embed_positions.weight"] = (npos, d)
    for i in range(cfg.decoder_layers):
        p = f"model.decoder.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        attn(p + ".encoder_attn")
        ln(p + ".encoder_attn_layer_norm")
        s[p + ".fc1.weight"] = (cfg.decoder_ffn_dim, d)
        s[p + ".fc1.bias"] = (cfg.decoder_ffn_dim,)
        s[p + ".fc2.weight"] = (d, cfg.decoder_ffn_dim)
        s[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")
    ln("model.decoder.layernorm_embedding")
    if pretraining:
        for name, din, ncls in (("mrm_head", d, cfg.num_labels), ("attribute_head", d, cfg.num_attributes),
                                ("relation_head", 2 * d, cfg.num_relations)):
            s[f"{name}.dense.weight"] = (d, din)
            s[f"{name}.dense.bias"] = (d,)
            s[f"{name}.out_proj.weight"] = (ncls, d)
            s[f"{name}.out_proj.bias"] = (ncls,)
    return s


def init_state_dict(cfg, seed=0, pretraining=False, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's rule (HF-3.0.2 PretrainedBartModel._init_weights):
    Linear/Embedding weights ~ N(0, init_std), biases 0, LayerNorm (1, 0), padding_idx row zero.
    The draw ORDER is this oracle's own (one generator, param_shapes order) — tests load the
    same tensors into the product model, they do not rely on RNG equality with the reference."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg, pretraining).items():
        if "layer_norm" in k or "layernorm" in k:
            sd[k] = torch.ones(shp, dtype=dtype) if k.endswith("weight") else torch.zeros(shp, dtype=dtype)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shp, dtype=dtype)
        else:
            sd[k] = (torch.randn(shp, generator=g) * cfg.init_std).to(dtype)
    sd["model.shared.weight"][cfg.pad_token_id].zero_()
    sd["model.encoder.embed_positions.weight"][cfg.pad_token_id].zero_()
    sd["model.decoder.embed_positions.weight"][cfg.pad_token_id].zero_()
    sd["final_logits_bias"] = torch.zeros(1, cfg.vocab_size, dtype=dtype)
    return sd


def full_state_dict(sd):
    """Add the tied aliases so the dict matches the reference's saved checkpoint keys."""
    out = dict(sd)
    out["model.encoder.embed_tokens.weight"] = sd["model.shared.weight"]
    out["model.decoder.embed_tokens.weight"] = sd["model.shared.weight"]
    return out


# ----------------------------------------------------------------------------- synthetic batches
def synthetic_batch(cfg, batch=16, n_regions=36, n_ctx=64, tgt_len=48, seed=1234, ragged=False):
    """SURVEY.md §8(d) synthetic VCG batch.  `ragged` right-pads 25 % of the rows (pad=1) and
    varies the region count, the edge cases the reference's collator produces
    (src/data/collation.py:68-213, src/data/tokenization.py:100-250)."""
    g = torch.Generator().manual_seed(seed)
    S_e = n_regions + n_ctx
    ids = torch.full((batch, S_e), cfg.pad_token_id, dtype=torch.long)
    mask = torch.zeros(batch, S_e, dtype=torch.long)
    feats = []
    for b in range(batch):
        R = n_regions
        n_txt = n_ctx - 5
        if ragged and b % 4 == 1:
            R = max(0, n_regions - 1 - int(torch.randint(0, max(1, n_regions // 2), (1,), generator=g)))
            n_txt = max(1, n_txt - int(torch.randint(1, max(2, n_txt // 2), (1,), generator=g)))
        txt = torch.randint(3, 50265, (n_txt,), generator=g)
        row = [50270, 50265] + [cfg.img_feat_id] * R + [50266, 50267] + txt.tolist() + [50268]
        ids[b, :len(row)] = torch.tensor(row)
        mask[b, :len(row)] = 1
        f = torch.relu(torch.randn(R, cfg.image_feature_size - 4, generator=g))
        x1 = torch.rand(R, generator=g) * 1024
        x2 = x1 + torch.rand(R, generator=g) * (1024 - x1)
        y1 = torch.rand(R, generator=g) * 768
        y2 = y1 + torch.rand(R, generator=g) * (768 - y1)
        box = torch.stack([x1, y1, x2, y2], 1)
        if R > 0:
            box[0] = torch.tensor([0.0, 0.0, 1024.0, 768.0])
        feats.append(torch.cat([f, box], 1).float())
    dec = torch.randint(3, 50265, (batch, tgt_len), generator=g)
    dec[:, 0] = cfg.bos_token_id
    labels = torch.randint(3, 50265, (batch, tgt_len), generator=g)
    labels[:, -1] = cfg.eos_token_id
    dmask = torch.ones(batch, tgt_len, dtype=torch.long)
    if ragged:
        for b in range(batch):
            if b % 4 == 2:
                n = int(torch.randint(2, tgt_len, (1,), generator=g))
                dec[b, n:] = cfg.pad_token_id
                labels[b, n - 1] = cfg.eos_token_id
                labels[b, n:] = -100
                dmask[b, n:] = 0
    return dict(input_ids=ids, attention_mask=mask, image_features=feats,
                decoder_input_ids=dec, decoder_attention_mask=dmask, labels=labels)


# ----------------------------------------------------------------------------- building blocks
def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _heads(x, h):  # [B,T,C] -> [B,h,T,dh]
    B, T, C = x.shape
    return x.view(B, T, h, C // h).transpose(1, 2)


def attention(sd, p, heads, query, key_states, key_padding_mask=None, attn_mask=None,
              cache: Optional[dict] = None, static_kv=False, p_drop=0.0, training=False):
    """HF-3.0.2 SelfAttention.forward in [B,T,C] layout.  `cache` is the per-layer dict entry
    {"prev_key","prev_value","prev_key_padding_mask"} keyed "self"/"encoder_decoder" by the
    caller; returns (out, new_cache_entry)."""
    B, Tq, C = query.shape
    dh = C // heads
    q = _heads(_lin(query, sd, p + ".q_proj") * dh ** -0.5, heads)
    if static_kv and cache is not None and "prev_key" in cache:
        k, v = cache["prev_key"], cache["prev_value"]
    else:
        src = key_states if static_kv else query
        k = _heads(_lin(src, sd, p + ".k_proj"), heads)
        v = _heads(_lin(src, sd, p + ".v_proj"), heads)
        if (not static_kv) and cache is not None and "prev_key" in cache:
            k = torch.cat([cache["prev_key"], k], 2)
            v = torch.cat([cache["prev_value"], v], 2)
    if cache is not None:
        prev_mask = cache.get("prev_key_padding_mask")
        if prev_mask is not None:
            key_padding_mask = prev_mask if static_kv else torch.cat([prev_mask, key_padding_mask], 1)
        elif key_padding_mask is not None:
            fill = torch.zeros(B, k.shape[2] - key_padding_mask.shape[1], dtype=key_padding_mask.dtype)
            key_padding_mask = torch.cat([fill, key_padding_mask], 1)
    new_cache = {"prev_key": k, "prev_value": v,
                 "prev_key_padding_mask": None if static_kv else key_padding_mask}
    w = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        w = w + attn_mask
    if key_padding_mask is not None:
        w = w.masked_fill(key_padding_mask[:, None, None, :], NEG_INF)
    w = F.softmax(w, dim=-1)
    w = F.dropout(w, p_drop, training)
    o = (w @ v).transpose(1, 2).reshape(B, Tq, C)
    return _lin(o, sd, p + ".out_proj"), new_cache


def _ffn(sd, p, x):
    return _lin(F.gelu(_lin(x, sd, p + ".fc1")), sd, p + ".fc2")


def embed_multimodal(sd, cfg, input_ids, image_features: List[torch.Tensor]):
    """src/model/modules.py:24-41 + :89-102 — Linear(2052->d) over the concatenated RoI rows,
    written over the token rows whose id is <img_feat> or <cls>."""
    emb = F.embedding(input_ids, sd["model.shared.weight"]).clone()
    slot = (input_ids == cfg.img_feat_id) | (input_ids == cfg.cls_token_id)
    for b, f in enumerate(image_features):
        if len(f) > 0:
            emb[b, slot[b]] = _lin(f, sd, "model.encoder.embed_images.linear")
    return emb


def encoder_forward(sd, cfg, input_ids, image_features, attention_mask=None, training=False):
    """src/model/modules.py:104-165; returns [B,S,C]."""
    pad = attention_mask.eq(0) if attention_mask is not None else None
    scale = math.sqrt(cfg.d_model) if cfg.scale_embedding else 1.0
    S = input_ids.shape[1]
    pos = sd["model.encoder.embed_positions.weight"][torch.arange(S) + cfg.extra_pos_embeddings]
    x = embed_multimodal(sd, cfg, input_ids, image_features) * scale + pos
    x = _ln(x, sd, "model.encoder.layernorm_embedding")
    x = F.dropout(x, cfg.dropout, training)
    for i in range(cfg.encoder_layers):
        p = f"model.encoder.layers.{i}"
        a, _ = attention(sd, p + ".self_attn", cfg.encoder_attention_heads, x, x, pad,
                         p_drop=cfg.attention_dropout, training=training)
        x = _ln(x + F.dropout(a, cfg.dropout, training), sd, p + ".self_attn_layer_norm")
        x = _ln(x + F.dropout(_ffn(sd, p, x), cfg.dropout, training), sd, p + ".final_layer_norm")
    return x


def prepare_decoder_inputs(cfg, input_ids, decoder_input_ids, decoder_attention_mask):
    """HF-3.0.2 _prepare_bart_decoder_inputs as called at src/model/model.py:63-70."""
    if decoder_input_ids is None:
        prev = input_ids.clone()
        last = (input_ids.ne(cfg.pad_token_id).sum(1) - 1).unsqueeze(-1)
        prev[:, 0] = input_ids.gather(1, last).squeeze(-1)
        prev[:, 1:] = input_ids[:, :-1]
        decoder_input_ids = prev
    T = decoder_input_ids.shape[1]
    if decoder_attention_mask is None:
        pad = decoder_input_ids.eq(cfg.pad_token_id)
        pad = pad if pad.any() else None
    else:
        pad = decoder_attention_mask.eq(0)
    causal = torch.triu(torch.full((T, T), NEG_INF), 1)
    return decoder_input_ids, pad, causal


def decoder_forward(sd, cfg, decoder_input_ids, enc_out, enc_attention_mask, dec_pad, causal,
                    caches=None, use_cache=False, training=False):
    """HF-3.0.2 BartDecoder.forward.  With use_cache only the last token is embedded, at
    position len-1 (+2 offset), and per-layer caches grow by one key/value."""
    enc_pad = enc_attention_mask.eq(0) if enc_attention_mask is not None else None
    T = decoder_input_ids.shape[1]
    off = cfg.extra_pos_embeddings
    if use_cache:
        pos = sd["model.decoder.embed_positions.weight"][torch.tensor([T - 1 + off])]
        ids = decoder_input_ids[:, -1:]
    else:
        pos = sd["model.decoder.embed_positions.weight"][torch.arange(T) + off]
        ids = decoder_input_ids
    scale = math.sqrt(cfg.d_model) if cfg.scale_embedding else 1.0
    x = F.embedding(ids, sd["model.shared.weight"]) * scale + pos
    x = _ln(x, sd, "model.decoder.layernorm_embedding")
    x = F.dropout(x, cfg.dropout, training)
    new_caches = []
    for i in range(cfg.decoder_layers):
        p = f"model.decoder.layers.{i}"
        lc = caches[i] if caches is not None else ({} if use_cache else None)
        a, c_self = attention(sd, p + ".self_attn", cfg.decoder_attention_heads, x, x, dec_pad, causal,
                              cache=(lc.get("self", {}) if lc is not None else None),
                              p_drop=cfg.attention_dropout, training=training)
        x = _ln(x + F.dropout(a, cfg.dropout, training), sd, p + ".self_attn_layer_norm")
        a, c_cross = attention(sd, p + ".encoder_attn", cfg.decoder_attention_heads, x, enc_out, enc_pad, None,
                               cache=(lc.get("encoder_decoder", {}) if lc is not None else None),
                               static_kv=True, p_drop=cfg.attention_dropout, training=training)
        x = _ln(x + F.dropout(a, cfg.dropout, training), sd, p + ".encoder_attn_layer_norm")
        x = _ln(x + F.dropout(_ffn(sd, p, x), cfg.dropout, training), sd, p + ".final_layer_norm")
        new_caches.append({"self": c_self, "encoder_decoder": c_cross})
    return x, (new_caches if use_cache else None)


def lm_logits(sd, h):
    return F.linear(h, sd["model.shared.weight"], sd["final_logits_bias"])


def forward_conditional_generation(sd, cfg, input_ids, image_features, attention_mask=None,
                                   decoder_input_ids=None, decoder_attention_mask=None, labels=None,
                                   training=False):
    """MultiModalBartForConditionalGeneration.forward without cache (src/model/model.py:325-405).
    Returns (loss or None, logits[B,T,V], decoder_hidden[B,T,C], encoder_out[B,S,C])."""
    dec_ids, dec_pad, causal = prepare_decoder_inputs(cfg, input_ids, decoder_input_ids, decoder_attention_mask)
    enc = encoder_forward(sd, cfg, input_ids, image_features, attention_mask, training)
    h, _ = decoder_forward(sd, cfg, dec_ids, enc, attention_mask, dec_pad, causal, training=training)
    logits = lm_logits(sd, h)
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits.view(-1, cfg.vocab_size), labels.view(-1), ignore_index=-100)
    return loss, logits, h, enc


def _cls_head(sd, name, x):
    return _lin(torch.tanh(_lin(x, sd, name + ".dense")), sd, name + ".out_proj")


def forward_pretraining(sd, cfg, input_ids, image_features, attention_mask=None, decoder_input_ids=None,
                        decoder_attention_mask=None, labels=None, mrm_labels=None, mrm_mask=None,
                        attribute_labels=None, attribute_mask=None, relation_labels=None, training=False):
    """MultiModalBartForPreTraining.forward (src/model/model.py:162-309): returns (losses, logits)."""
    if mrm_labels is not None and mrm_mask is None:
        raise ValueError('"mrm_mask" cannot be None while "mrm_labels" is set')
    dec_ids, dec_pad, causal = prepare_decoder_inputs(cfg, input_ids, decoder_input_ids, decoder_attention_mask)
    enc = encoder_forward(sd, cfg, input_ids, image_features, attention_mask, training)
    h, _ = decoder_forward(sd, cfg, dec_ids, enc, attention_mask, dec_pad, causal, training=training)
    losses = {}
    mrm_loss = attr_loss = rel_loss = lm_loss = 0
    if mrm_labels is not None:
        rep = h[mrm_mask.bool()]
        if len(rep) > 0:
            lp = F.log_softmax(_cls_head(sd, "mrm_head", rep), dim=1)
            mrm_loss = F.kl_div(lp, torch.cat(mrm_labels, 0), reduction="batchmean") * cfg.mrm_loss_factor
            losses["mrm_loss"] = mrm_loss
    if attribute_labels is not None:
        rep = h[attribute_mask.bool()]
        if len(rep) > 0:
            attr_loss = F.cross_entropy(_cls_head(sd, "attribute_head", rep),
                                        torch.cat(attribute_labels, 0).reshape(-1)) * cfg.attribute_loss_factor
            losses["attribute_loss"] = attr_loss
    if relation_labels is not None:
        pairs, ids = [], []
        for b, rels in enumerate(relation_labels):
            for r in rels:
                pairs.append(torch.cat([h[b][r["object_index"]], h[b][r["subject_index"]]], 0))
                ids.append(r["label"])
        if pairs:
            rel_loss = F.cross_entropy(_cls_head(sd, "relation_head", torch.stack(pairs)),
                                       torch.tensor(ids, dtype=torch.long)) * cfg.relation_loss_factor
            losses["relation_loss"] = rel_loss
    logits = lm_logits(sd, h)
    if labels is not None:
        lab = labels.clone()
        lab[lab == cfg.cls_token_id] = -100
        lm_loss = F.cross_entropy(logits.view(-1, cfg.vocab_size), lab.reshape(-1), ignore_index=-100) * cfg.lm_loss_factor
        losses["lm_loss"] = lm_loss
    if labels is not None or mrm_labels is not None or attribute_labels is not None or relation_labels is not None:
        losses["loss"] = lm_loss + mrm_loss + attr_loss + rel_loss
    return losses, logits


# ----------------------------------------------------------------------------- generation
def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, min_tokens_to_keep=1):
    """HF-3.0.2 generation_utils.top_k_top_p_filtering (in place on `logits`)."""
    if top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits[logits < kth] = NEG_INF
    if top_p < 1.0:
        s_logits, s_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(s_logits, dim=-1), dim=-1)
        rm = cum > top_p
        if min_tokens_to_keep > 1:
            rm[..., :min_tokens_to_keep] = 0
        rm[..., 1:] = rm[..., :-1].clone()
        rm[..., 0] = 0
        logits[rm.scatter(1, s_idx, rm)] = NEG_INF
    return logits


class _BeamHyps:
    """HF-3.0.2 BeamHypotheses."""

    def __init__(self, num_beams, max_length, length_penalty, early_stopping):
        self.max_length = max_length - 1
        self.length_penalty, self.early_stopping, self.num_beams = length_penalty, early_stopping, num_beams
        self.beams, self.worst_score = [], 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                ranked = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


def _postprocess_scores(scores, input_ids, cur_len, min_length, eos_token_id, repetition_penalty):
    """HF-3.0.2 postprocess_next_token_scores for the knobs the reference's callers reach
    (repetition_penalty, min_length); n-gram / bad-word bans are not used on this path
    (src/generation.py:22-32, vcg_train.py:190) and raise in generate()."""
    if repetition_penalty != 1.0:
        for i in range(scores.shape[0]):
            for tok in set(input_ids[i].tolist()):
                scores[i, tok] = scores[i, tok] * repetition_penalty if scores[i, tok] < 0 else scores[i, tok] / repetition_penalty
    if eos_token_id is not None and cur_len < min_length:
        scores[:, eos_token_id] = NEG_INF
    return scores


def _reorder_caches(caches, beam_idx):
    """src/model/mixins.py:419-434 (+ HF-3.0.2 _reorder_buffer)."""
    out = []
    for lc in caches:
        out.append({k: {n: (t.index_select(0, beam_idx) if t is not None else None) for n, t in c.items()}
                    for k, c in lc.items()})
    return out


@torch.no_grad()
def generate(sd, cfg, input_ids, image_features, attention_mask=None, max_length=None, min_length=None,
             do_sample=None, early_stopping=None, num_beams=None, temperature=None, top_k=None, top_p=None,
             repetition_penalty=None, length_penalty=None, num_return_sequences=None, use_cache=True,
             no_repeat_ngram_size=0, bad_words_ids=None, generator=None):
    """GenerationMixin.generate (src/model/mixins.py:33-384) + the HF-3.0.2 greedy / sampling /
    beam loops it dispatches to.  `generator` seeds torch.multinomial for sampling."""
    def dflt(v, name):
        return v if v is not None else getattr(cfg, name)
    max_length, min_length = dflt(max_length, "max_length"), dflt(min_length, "min_length")
    do_sample, early_stopping = dflt(do_sample, "do_sample"), dflt(early_stopping, "early_stopping")
    num_beams, temperature = dflt(num_beams, "num_beams"), dflt(temperature, "temperature")
    top_k, top_p = dflt(top_k, "top_k"), dflt(top_p, "top_p")
    repetition_penalty, length_penalty = dflt(repetition_penalty, "repetition_penalty"), dflt(length_penalty, "length_penalty")
    num_return_sequences = dflt(num_return_sequences, "num_return_sequences")
    if no_repeat_ngram_size or bad_words_ids:
        raise NotImplementedError("n-gram / bad-word bans are outside the reference's call sites")
    pad, eos, V = cfg.pad_token_id, cfg.eos_token_id, cfg.vocab_size
    B = input_ids.shape[0]
    if not do_sample:
        assert num_return_sequences == 1 if num_beams == 1 else num_beams >= num_return_sequences
    if attention_mask is None:
        attention_mask = input_ids.ne(pad).long() if (input_ids == pad).any() else torch.ones_like(input_ids)
    eff_bs, eff_mult = (B * num_return_sequences, num_return_sequences) if do_sample else (B, 1)
    enc = encoder_forward(sd, cfg, input_ids, image_features, attention_mask)
    rows = eff_bs * num_beams
    if num_return_sequences > 1 or num_beams > 1:
        exp = torch.arange(B).view(-1, 1).repeat(1, num_beams * eff_mult).view(-1)
        attention_mask = attention_mask.index_select(0, exp)
        enc = enc.index_select(0, exp)
    seq = torch.full((rows, 1), cfg.decoder_start_token_id, dtype=torch.long)
    cur_len = 1
    assert cur_len < max_length
    caches = None

    def step(seq, caches):
        if use_cache:
            h, caches = decoder_forward(sd, cfg, seq, enc, attention_mask, None, None, caches, use_cache=True)
        else:
            ids, dpad, causal = prepare_decoder_inputs(cfg, None, seq, None)
            h, _ = decoder_forward(sd, cfg, ids, enc, attention_mask, dpad, causal)
        return lm_logits(sd, h)[:, -1, :], caches

    if num_beams == 1:
        unfinished = torch.ones(rows, dtype=torch.long)
        sent_len = torch.full((rows,), max_length, dtype=torch.long)
        while cur_len < max_length:
            logits, caches = step(seq, caches)
            scores = _postprocess_scores(logits, seq, cur_len, min_length, eos, repetition_penalty)
            if do_sample:
                if temperature != 1.0:
                    scores = scores / temperature
                probs = F.softmax(top_k_top_p_filtering(scores, top_k, top_p), dim=-1)
                nxt = torch.multinomial(probs, 1, generator=generator).squeeze(1)
            else:
                nxt = torch.argmax(logits, dim=-1)
            tok = nxt * unfinished + pad * (1 - unfinished)
            seq = torch.cat([seq, tok.unsqueeze(-1)], -1)
            cur_len += 1
            is_eos = tok == eos
            sent_len.masked_fill_((unfinished * is_eos.long()).bool(), cur_len)
            unfinished = unfinished * (~is_eos).long()
            if unfinished.max() == 0:
                break
        if sent_len.min().item() != sent_len.max().item():
            out = torch.full((rows, sent_len.max().item()), pad, dtype=torch.long)
        else:
            out = seq
        for i in range(rows):
            out[i, :sent_len[i]] = seq[i, :sent_len[i]]
        return out

    # ---- beam search
    hyps = [_BeamHyps(num_beams, max_length, length_penalty, early_stopping) for _ in range(eff_bs)]
    beam_scores = torch.zeros(eff_bs, num_beams)
    if not do_sample:
        beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    done = [False] * eff_bs
    while cur_len < max_length:
        logits, caches = step(seq, caches)
        if not do_sample:  # adjust_logits_during_generation, src/model/mixins.py:400-417
            if cur_len == 1:
                keep = logits[:, cfg.bos_token_id].clone()
                logits.fill_(NEG_INF)
                logits[:, cfg.bos_token_id] = keep
            if cur_len == max_length - 1 and eos is not None:
                keep = logits[:, eos].clone()
                logits.fill_(NEG_INF)
                logits[:, eos] = keep
        scores = F.log_softmax(logits, dim=-1)
        scores = _postprocess_scores(scores, seq, cur_len, min_length, eos, repetition_penalty)
        if do_sample:
            _s = scores + beam_scores[:, None]
            if temperature != 1.0:
                _s = _s / temperature
            _s = top_k_top_p_filtering(_s, top_k, top_p, min_tokens_to_keep=2).contiguous().view(eff_bs, num_beams * V)
            nt = torch.multinomial(F.softmax(_s, dim=-1), 2 * num_beams, generator=generator)
            ns = torch.gather(_s, -1, nt)
            ns, order = torch.sort(ns, descending=True, dim=1)
            nt = torch.gather(nt, -1, order)
        else:
            ns = (scores + beam_scores[:, None]).view(eff_bs, num_beams * V)
            ns, nt = torch.topk(ns, 2 * num_beams, dim=1, largest=True, sorted=True)
        nxt_beam = []
        for b in range(eff_bs):
            if done[b]:
                nxt_beam.extend([(0, pad, 0)] * num_beams)
                continue
            sent = []
            for rank, (tid, sc) in enumerate(zip(nt[b].tolist(), ns[b].tolist())):
                beam_id, tok = tid // V, tid % V
                row = b * num_beams + beam_id
                if eos is not None and tok == eos:
                    if rank >= num_beams:
                        continue
                    hyps[b].add(seq[row].clone(), sc)
                else:
                    sent.append((sc, tok, row))
                if len(sent) == num_beams:
                    break
            done[b] = done[b] or hyps[b].is_done(ns[b].max().item(), cur_len)
            assert len(sent) == num_beams, "Beam should always be full"
            nxt_beam.extend(sent)
        if all(done):
            break
        beam_scores = torch.tensor([x[0] for x in nxt_beam], dtype=torch.float)
        beam_tok = torch.tensor([x[1] for x in nxt_beam], dtype=torch.long)
        beam_idx = torch.tensor([x[2] for x in nxt_beam], dtype=torch.long)
        seq = torch.cat([seq[beam_idx], beam_tok.unsqueeze(1)], -1)
        cur_len += 1
        if caches is not None:
            caches = _reorder_caches(caches, beam_idx)
            # the reference also index_selects the (already expanded) encoder output and mask
            enc = enc.index_select(0, beam_idx)
            attention_mask = attention_mask.index_select(0, beam_idx)
    for b in range(eff_bs):
        if done[b]:
            continue
        for j in range(num_beams):
            row = b * num_beams + j
            hyps[b].add(seq[row], beam_scores[row].item())
    n_out_per = 1 if do_sample else num_return_sequences
    out_bs = eff_bs if do_sample else eff_bs * num_return_sequences
    best, lens = [], []
    for h in hyps:
        ranked = sorted(h.beams, key=lambda x: x[0])
        for _ in range(n_out_per):
            hyp = ranked.pop()[1]
            best.append(hyp)
            lens.append(len(hyp))
    if min(lens) != max(lens):
        width = min(max(lens) + 1, max_length)
        out = torch.full((out_bs, width), pad, dtype=torch.long)
        for i, hyp in enumerate(best):
            out[i, :lens[i]] = hyp
            if lens[i] < max_length:
                out[i, lens[i]] = eos
    else:
        out = torch.stack(best).long()
    return out


# ----------------------------------------------------------------------------- optimizer
def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-5, betas=(0.9, 0.999), eps=1e-6,
               weight_decay=0.0, correct_bias=True):
    """HF-3.0.2 transformers.AdamW.step for one step number `step` (1-based), in place.
    eps is added to the un-corrected sqrt(v); bias correction folds into the step size;
    decoupled weight decay is applied after the Adam update."""
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = v.sqrt().add_(eps)
        step_size = lr
        if correct_bias:
            step_size = step_size * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
        p.addcdiv_(m, denom, value=-step_size)
        if weight_decay > 0.0:
            p.add_(p, alpha=-lr * weight_decay)
