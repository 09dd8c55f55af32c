"""Synthetic batches of the benchmark workloads — used by bench.py and the timing scripts.

They reproduce the tensors the reference's collator hands to the model (src/data/collation.py:68-213,
src/data/tokenization.py:100-250): `<intent> <img> <img_feat>xR </img> <event> text </event>` on the encoder
side, `<s> text` / `text </s>` on the decoder side, RoI rows = relu(N(0,1))[2048] ++ pixel box [4] with the
first box covering the whole image (scripts/prepare_vcg.py:34), and for pre-training the masked-region /
attribute / relation label structures of src/model/model.py:162-309.

`cfg` is anything with pad_token_id / bos_token_id / eos_token_id / img_feat_id / cls_token_id /
image_feature_size (+ num_labels / num_attributes / num_relations for pre-training).  tests/test_boundary.py
checks that these generators and the oracle's own (oracle/kmbart_oracle.py) emit identical tensors."""
import torch

TXT_LO, TXT_HI = 3, 50265            # plain BPE ids; 50265.. are the task tokens of src/data/tokenization.py:36-57
TOK_IMG, TOK_IMG_END, TOK_EVENT, TOK_EVENT_END, TOK_INTENT = 50265, 50266, 50267, 50268, 50270


def _regions(cfg, R, g):
    f = torch.relu(torch.randn(R, cfg.image_feature_size - 4, generator=g))
    x1 = torch.rand(R, generator=g) * 1024
    x2 = x1 + torch.rand(R, generator=g) * (1024 - x1)
    y1 = torch.rand(R, generator=g) * 768
    y2 = y1 + torch.rand(R, generator=g) * (768 - y1)
    box = torch.stack([x1, y1, x2, y2], 1)
    if R > 0:
        box[0] = torch.tensor([0.0, 0.0, 1024.0, 768.0])
    return torch.cat([f, box], 1).float()


def synthetic_batch(cfg, batch=16, n_regions=36, n_ctx=64, tgt_len=48, seed=1234, ragged=False):
    """VCG fine-tuning batch.  `ragged`: every 4th row has fewer regions / a shorter context (right padding,
    pad = 1) and every 4th target is cut short (labels -100 beyond its </s>)."""
    g = torch.Generator().manual_seed(seed)
    S_e = n_regions + n_ctx
    ids = torch.full((batch, S_e), cfg.pad_token_id, dtype=torch.long)
    mask = torch.zeros(batch, S_e, dtype=torch.long)
    feats = []
    for b in range(batch):
        R, n_txt = n_regions, n_ctx - 5
        if ragged and b % 4 == 1:
            R = max(0, n_regions - 1 - int(torch.randint(0, max(1, n_regions // 2), (1,), generator=g)))
            n_txt = max(1, n_txt - int(torch.randint(1, max(2, n_txt // 2), (1,), generator=g)))
        txt = torch.randint(TXT_LO, TXT_HI, (n_txt,), generator=g)
        row = [TOK_INTENT, TOK_IMG] + [cfg.img_feat_id] * R + [TOK_IMG_END, TOK_EVENT] + txt.tolist() + [TOK_EVENT_END]
        ids[b, :len(row)] = torch.tensor(row)
        mask[b, :len(row)] = 1
        feats.append(_regions(cfg, R, g))
    dec = torch.randint(TXT_LO, TXT_HI, (batch, tgt_len), generator=g)
    dec[:, 0] = cfg.bos_token_id
    labels = torch.randint(TXT_LO, TXT_HI, (batch, tgt_len), generator=g)
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


def synthetic_pretrain_batch(cfg, batch=128, n_regions=36, n_ctx=64, tgt_len=48, seed=1234, mrm_p=0.2, n_attr=16, n_rel=32):
    """Multitask pre-training batch (BASELINE configs[2]): decoder input `<img> slots </img> <s> text`
    (S_d = R + 2 + tgt_len), masked regions replaced by <cls> with their feature zeroed and box kept
    (src/data/collation.py:113-132), soft region labels, attribute labels on unmasked slots and relation triples
    as host dicts (src/training.py:45)."""
    b0 = synthetic_batch(cfg, batch=batch, n_regions=n_regions, n_ctx=n_ctx, tgt_len=tgt_len, seed=seed)
    g = torch.Generator().manual_seed(seed + 9)
    B, R, T = batch, n_regions, tgt_len
    Sd = R + 2 + T
    dec = torch.full((B, Sd), cfg.pad_token_id, dtype=torch.long)
    labels = torch.full((B, Sd), -100, dtype=torch.long)
    mrm_mask = torch.zeros(B, Sd, dtype=torch.bool)
    attr_mask = torch.zeros(B, Sd, dtype=torch.bool)
    mrm_labels, attr_labels, rel_labels = [], [], []
    for b in range(B):
        slots = [cfg.img_feat_id] * R
        masked = [i for i in range(R) if torch.rand(1, generator=g).item() < mrm_p]
        for i in masked:
            slots[i] = cfg.cls_token_id
            mrm_mask[b, 1 + i] = True
            b0["image_features"][b][i, :cfg.image_feature_size - 4] = 0
        dec[b] = torch.tensor([TOK_IMG] + slots + [TOK_IMG_END, cfg.bos_token_id] + b0["decoder_input_ids"][b, 1:T].tolist())
        labels[b, :R + 2] = cfg.cls_token_id          # turned into -100 by the model (src/model/model.py:297-298)
        labels[b, R + 2:] = b0["labels"][b, :T]
        mrm_labels.append(torch.softmax(torch.randn(len(masked), cfg.num_labels, generator=g), -1))
        attr_slots = [i for i in range(R) if i not in masked][:n_attr]
        for i in attr_slots:
            attr_mask[b, 1 + i] = True
        attr_labels.append(torch.randint(0, cfg.num_attributes, (len(attr_slots),), generator=g))
        rels = []
        for _ in range(n_rel):
            o, s = torch.randint(0, R, (2,), generator=g).tolist()
            rels.append({"object_index": 1 + o, "subject_index": 1 + s,
                         "label": int(torch.randint(0, cfg.num_relations, (1,), generator=g))})
        rel_labels.append(rels)
    b0.update(decoder_input_ids=dec, labels=labels, decoder_attention_mask=torch.ones(B, Sd, dtype=torch.long),
              mrm_labels=mrm_labels, mrm_mask=mrm_mask, attribute_labels=attr_labels, attribute_mask=attr_mask,
              relation_labels=rel_labels)
    return b0


def to_device(batch, device, pin=False):
    """Host batch -> device (relation_labels stay host dicts, like the reference's loop)."""
    out = {}
    for k, v in batch.items():
        if k == "relation_labels":
            out[k] = v
        elif isinstance(v, list):
            out[k] = [(t.pin_memory() if pin else t).to(device) if device is not None else (t.pin_memory() if pin else t) for t in v]
        else:
            out[k] = (v.pin_memory() if pin else v).to(device) if device is not None else (v.pin_memory() if pin else v)
    return out
