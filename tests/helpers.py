"""Shared helpers for the GPU parity tests (product model <-> CPU oracle)."""
import torch

from oracle import kmbart_oracle as O


def small_config(**kw):
    """2+2 layer, d=128 (2 heads x 64) model: the oracle finishes in well under a second."""
    d = dict(d_model=128, encoder_layers=2, decoder_layers=2, encoder_attention_heads=2, decoder_attention_heads=2,
             encoder_ffn_dim=256, decoder_ffn_dim=256, dropout=0.0, max_position_embeddings=256)
    d.update(kw)
    return O.OracleConfig(**d)


def product_config(ocfg):
    from src.model.config import MultiModalBartConfig
    keys = ["vocab_size", "d_model", "image_feature_size", "encoder_layers", "decoder_layers", "encoder_attention_heads",
            "decoder_attention_heads", "encoder_ffn_dim", "decoder_ffn_dim", "max_position_embeddings",
            "extra_pos_embeddings", "dropout", "attention_dropout", "activation_dropout", "init_std", "pad_token_id",
            "bos_token_id", "eos_token_id", "decoder_start_token_id", "img_feat_id", "cls_token_id", "num_labels",
            "num_attributes", "num_relations", "lm_loss_factor", "mrm_loss_factor", "attribute_loss_factor",
            "relation_loss_factor"]
    return MultiModalBartConfig(**{k: getattr(ocfg, k) for k in keys})


def load_oracle_weights(model, sd):
    missing, unexpected = model.load_state_dict(O.full_state_dict(sd), strict=False)
    assert not unexpected, unexpected
    assert not missing, missing


def to_cuda_batch(batch, device="cuda"):
    out = {}
    for k, v in batch.items():
        if isinstance(v, list):
            out[k] = [t.to(device) for t in v]
        else:
            out[k] = v.to(device)
    return out


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()
