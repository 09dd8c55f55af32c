"""Seeded inputs of the golden cases — shared by tests/golden/make_golden.py (which runs the
reference on them) and by the tests (which run the oracle / the CUDA path on them)."""
import torch

from oracle import kmbart_oracle as O

LOGIT_COLS = torch.tensor([0, 1, 2, 3, 17, 100, 999, 4242, 20000, 31337, 50264, 50265, 50273, 50276, 50300, 50319])
GRAD_SLICE_NAMES = ["model.encoder.embed_images.linear.weight", "model.encoder.layers.0.fc1.weight",
                    "model.decoder.layers.1.encoder_attn.v_proj.weight", "model.decoder.layernorm_embedding.weight",
                    "model.encoder.embed_positions.weight"]
SAMPLE_SEED = 77
GENERATE_CASES = {
    "greedy": dict(max_length=8),
    "greedy_min_len": dict(max_length=7, min_length=7),
    "beam3_early": dict(max_length=8, num_beams=3, early_stopping=True),
    "beam4_ret2": dict(max_length=9, num_beams=4, num_return_sequences=2),
    "beam2_lenpen": dict(max_length=8, num_beams=2, length_penalty=2.0, repetition_penalty=1.3),
    "sample_topk": dict(max_length=8, do_sample=True, top_k=10),
    "sample_topp_ret2": dict(max_length=7, do_sample=True, top_k=0, top_p=0.8, num_return_sequences=2, temperature=0.7),
}
ADAMW = dict(lr=1e-3, weight_decay=0.01)

_CFG_KEYS = ["vocab_size", "d_model", "image_feature_size", "encoder_layers", "decoder_layers", "encoder_attention_heads",
             "decoder_attention_heads", "encoder_ffn_dim", "decoder_ffn_dim", "max_position_embeddings", "dropout",
             "init_std", "num_labels", "num_attributes", "num_relations", "lm_loss_factor", "mrm_loss_factor",
             "attribute_loss_factor", "relation_loss_factor"]


def config_kwargs(ocfg):
    return {k: getattr(ocfg, k) for k in _CFG_KEYS}


def small_config(**kw):
    d = dict(d_model=128, encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             encoder_ffn_dim=256, decoder_ffn_dim=256, dropout=0.0, max_position_embeddings=256)
    d.update(kw)
    return O.OracleConfig(**d)


def perturb(sd, seed=5):
    """Non-trivial biases / LayerNorm parameters / final_logits_bias so every term is exercised."""
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k == "final_logits_bias":
            continue
        if k.endswith(".bias") or "layer_norm" in k or "layernorm" in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    sd["final_logits_bias"] = 0.1 * torch.randn(sd["final_logits_bias"].shape, generator=g)
    return sd


def checksum(sd, batch):
    s = sum(float(v.double().abs().sum()) for v in sd.values())
    b = float(batch["input_ids"].double().sum()) + sum(float(f.double().abs().sum()) for f in batch["image_features"])
    return torch.tensor([s, b], dtype=torch.float64)


def case_forward():
    ocfg = small_config()
    sd = perturb(O.init_state_dict(ocfg, seed=0))
    batch = O.synthetic_batch(ocfg, batch=4, n_regions=6, n_ctx=14, tgt_len=10, seed=3, ragged=True)
    return ocfg, sd, batch


def case_pretrain():
    """config/pretrain_base.json loss factors and head sizes on the small trunk; decoder input is
    <img> slots </img> <s> text (src/data/tokenization.py:197-250), heads read decoder states at slots."""
    ocfg = small_config(num_labels=1601, num_attributes=129, num_relations=129, lm_loss_factor=5.0,
                        mrm_loss_factor=1.0, attribute_loss_factor=1.0, relation_loss_factor=1.0)
    sd = perturb(O.init_state_dict(ocfg, seed=1, pretraining=True), seed=6)
    B, R, T = 4, 6, 9
    batch = O.synthetic_batch(ocfg, batch=B, n_regions=R, n_ctx=14, tgt_len=T, seed=4, ragged=False)
    g = torch.Generator().manual_seed(9)
    Sd = R + 2 + T
    dec = torch.full((B, Sd), ocfg.pad_token_id, dtype=torch.long)
    labels = torch.full((B, Sd), -100, dtype=torch.long)
    mrm_mask = torch.zeros(B, Sd, dtype=torch.bool)
    attr_mask = torch.zeros(B, Sd, dtype=torch.bool)
    mrm_labels, attr_labels, rel_labels = [], [], []
    for b in range(B):
        slots = [ocfg.img_feat_id] * R
        masked = [i for i in range(R) if (i + b) % 3 == 0]
        for i in masked:
            slots[i] = ocfg.cls_token_id
            mrm_mask[b, 1 + i] = True
        dec[b] = torch.tensor([50265] + slots + [50266, ocfg.bos_token_id] + batch["decoder_input_ids"][b, 1:T].tolist())
        labels[b, :R + 2] = torch.tensor([ocfg.cls_token_id] * (R + 2))     # turned into -100 by the model (:297-298)
        labels[b, R + 2:] = batch["labels"][b, :T]
        mrm_labels.append(torch.softmax(torch.randn(len(masked), ocfg.num_labels, generator=g), -1))
        attr_slots = [i for i in range(R) if i not in masked][:3]
        for i in attr_slots:
            attr_mask[b, 1 + i] = True
        attr_labels.append(torch.randint(0, ocfg.num_attributes, (len(attr_slots),), generator=g))
        rels = []
        for _ in range(b + 1):
            o, s = torch.randint(0, R, (2,), generator=g).tolist()
            rels.append({"object_index": 1 + o, "subject_index": 1 + s,
                         "label": int(torch.randint(0, ocfg.num_relations, (1,), generator=g))})
        rel_labels.append(rels)
    batch.update(decoder_input_ids=dec, labels=labels, decoder_attention_mask=torch.ones(B, Sd, dtype=torch.long),
                 mrm_labels=mrm_labels, mrm_mask=mrm_mask, attribute_labels=attr_labels, attribute_mask=attr_mask,
                 relation_labels=rel_labels)
    return ocfg, sd, batch


def case_adamw():
    g = torch.Generator().manual_seed(21)
    shapes = [(37, 5), (130,), (9000,), (3,)]
    params = [torch.randn(s, generator=g) for s in shapes]
    grads_seq = [[torch.randn(s, generator=g) * (0.1 + i) for s in shapes] for i in range(3)]
    return params, grads_seq
