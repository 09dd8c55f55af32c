"""Parameter-holding modules of the KM-BART encoder/decoder — drop-in for the reference's
src/model/modules.py (ImageEmbedding :19-41, MultiModalBartEncoder :49-165) and for the
HF-3.0.2 modeling_bart classes it imports (EncoderLayer, DecoderLayer, SelfAttention,
LearnedPositionalEmbedding, LayerNorm, BartDecoder, BartClassificationHead).

Module / parameter names and registration order reproduce the reference so that
state_dict(), parameters() order (optimizer-state checkpoints) and DDP see the same model.
The arithmetic is NOT here: these modules are containers; the forward/backward of the whole
stack is executed by kmbart.engine.Engine through hand-written sm_100a kernels.  A submodule
called on its own (other than the encoder, which generate()/sample_sentence() call directly)
raises instead of silently running a PyTorch fallback."""
import math
import weakref

import torch
import torch.nn as nn


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True):
    return nn.LayerNorm(normalized_shape, eps, elementwise_affine)


class _KernelOnly(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__} is a parameter container; its arithmetic runs inside the fused "
            "sm_100a engine of the owning MultiModalBart* model (no standalone PyTorch path).")


class LearnedPositionalEmbedding(nn.Embedding):
    """Same storage as HF-3.0.2: num_embeddings + offset rows, lookup at arange(seq_len) + offset."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx, offset):
        self.offset = offset
        assert padding_idx is not None
        super().__init__(num_embeddings + offset, embedding_dim, padding_idx=padding_idx)


class SelfAttention(_KernelOnly):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, encoder_decoder_attention=False):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.encoder_decoder_attention = encoder_decoder_attention
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.cache_key = "encoder_decoder" if encoder_decoder_attention else "self"


class EncoderLayer(_KernelOnly):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = SelfAttention(self.embed_dim, config.encoder_attention_heads, dropout=config.attention_dropout)
        self.normalize_before = config.normalize_before
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.dropout = config.dropout
        self.activation_dropout = config.activation_dropout
        self.fc1 = nn.Linear(self.embed_dim, config.encoder_ffn_dim)
        self.fc2 = nn.Linear(config.encoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)


class DecoderLayer(_KernelOnly):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = SelfAttention(self.embed_dim, config.decoder_attention_heads, dropout=config.attention_dropout)
        self.dropout = config.dropout
        self.activation_dropout = config.activation_dropout
        self.normalize_before = config.normalize_before
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.encoder_attn = SelfAttention(self.embed_dim, config.decoder_attention_heads, dropout=config.attention_dropout,
                                          encoder_decoder_attention=True)
        self.encoder_attn_layer_norm = LayerNorm(self.embed_dim)
        self.fc1 = nn.Linear(self.embed_dim, config.decoder_ffn_dim)
        self.fc2 = nn.Linear(config.decoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)


class BartDecoder(_KernelOnly):
    def __init__(self, config, embed_tokens):
        super().__init__()
        self.dropout = config.dropout
        self.layerdrop = config.decoder_layerdrop
        self.padding_idx = embed_tokens.padding_idx
        self.max_target_positions = config.max_position_embeddings
        self.embed_scale = math.sqrt(config.d_model) if config.scale_embedding else 1.0
        self.embed_tokens = embed_tokens
        self.embed_positions = LearnedPositionalEmbedding(config.max_position_embeddings, config.d_model,
                                                          self.padding_idx, config.extra_pos_embeddings)
        self.layers = nn.ModuleList([DecoderLayer(config) for _ in range(config.decoder_layers)])
        self.layernorm_embedding = LayerNorm(config.d_model) if config.normalize_embedding else nn.Identity()
        self.layer_norm = None


class BartClassificationHead(_KernelOnly):
    """dropout -> dense -> tanh -> dropout -> out_proj (HF-3.0.2), used by the pretraining heads."""

    def __init__(self, input_dim, inner_dim, num_classes, pooler_dropout):
        super().__init__()
        self.dense = nn.Linear(input_dim, inner_dim)
        self.dropout = nn.Dropout(p=pooler_dropout)
        self.out_proj = nn.Linear(inner_dim, num_classes)


class ImageEmbedding(_KernelOnly):
    def __init__(self, image_dim, final_dim):
        super().__init__()
        self.linear = nn.Linear(image_dim, final_dim)


class MultiModalBartEncoder(nn.Module):
    """Encoder container.  Calling it runs the fused encoder plan of the owning model — this is the
    entry generate() and src/model/utils.py:sample_sentence use (`model.get_encoder()(...)`)."""

    def __init__(self, config, embed_tokens):
        super().__init__()
        self.img_feat_id = config.img_feat_id
        self.cls_token_id = config.cls_token_id
        self.dropout = config.dropout
        self.layerdrop = config.encoder_layerdrop
        self.indentity = nn.Identity()   # (sic) attribute kept for parity with the reference
        embed_dim = embed_tokens.embedding_dim
        self.embed_scale = math.sqrt(embed_dim) if config.scale_embedding else 1.0
        self.padding_idx = embed_tokens.padding_idx
        self.max_source_positions = config.max_position_embeddings
        self.embed_tokens = embed_tokens
        self.embed_images = ImageEmbedding(config.image_feature_size, embed_dim)
        if config.static_position_embeddings:
            raise NotImplementedError("static (sinusoidal) position embeddings are not on the KM-BART path")
        self.embed_positions = LearnedPositionalEmbedding(config.max_position_embeddings, embed_dim, self.padding_idx,
                                                          config.extra_pos_embeddings)
        self.layers = nn.ModuleList([EncoderLayer(config) for _ in range(config.encoder_layers)])
        self.layernorm_embedding = LayerNorm(embed_dim) if config.normalize_embedding else nn.Identity()
        self.layer_norm = None
        if config.normalize_before or config.add_final_layer_norm or not config.normalize_embedding:
            raise NotImplementedError("only the post-LN BART layout of config/vcg_base.json / pretrain_base.json "
                                      "(normalize_before=False, normalize_embedding=True) is implemented")
        self._owner_ref = None

    def _owner(self):
        o = self._owner_ref() if self._owner_ref is not None else None
        if o is None:
            raise RuntimeError("encoder is not attached to a MultiModalBart* model")
        return o

    def forward(self, input_ids, image_features, attention_mask=None, output_attentions=False, output_hidden_states=False):
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / per-layer states are never materialised by the fused kernels")
        owner = self._owner()
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("stand-alone encoder calls are inference-only; train through model.forward(labels=...)")
        if owner.precision == "fp32":
            return owner._fp32().encoder(input_ids, image_features, attention_mask), [], []
        enc, _, _ = owner._engine().infer_forward(input_ids, image_features, attention_mask, None, None, encoder_only=True)
        return enc.clone(), [], []
