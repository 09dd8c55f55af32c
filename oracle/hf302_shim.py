"""Stand-in for the `transformers==3.0.2` symbols that fomalhautb/KM-BART imports — TEST
INFRASTRUCTURE, NOT PRODUCT CODE (same rule as oracle/kmbart_oracle.py).

Why it exists: the reference's own model code (`/root/reference/src/model/*.py`) is pure Python
but imports `transformers.modeling_bart`, `transformers.generation_utils` and a few names of
`transformers.modeling_utils` that only exist in transformers 3.0.2 (pinned at
environment.yaml:159).  That version is not installable here (no network; the image has 5.5),
so the reference cannot be imported unmodified.  `install()` registers small modules under
those three names in `sys.modules`; after that `import src.model.model` from
`/root/reference` succeeds and the reference's OWN forward / generate / from_pretrained code
runs on CPU.  tests/golden/make_golden.py uses this to produce the committed golden vectors,
and tests/test_oracle.py (CPU suite, only where /root/reference exists) re-checks the oracle
against a live run.

What is restated here (from the published transformers 3.0.2 sources, written as nn.Modules so
the reference's classes can subclass / instantiate them; import sites in parentheses):
  modeling_bart: PretrainedBartModel, BartDecoder, DecoderLayer, EncoderLayer, SelfAttention,
    LearnedPositionalEmbedding, SinusoidalPositionalEmbedding, LayerNorm, BartClassificationHead,
    invert_mask, _prepare_bart_decoder_inputs, _make_linear_from_emb, _filter_out_falsey_values,
    _reorder_buffer                      (src/model/model.py:8-15, modules.py:8-14, mixins.py:11-14)
  generation_utils: logger, Iterable, top_k_top_p_filtering, BeamHypotheses and the
    _generate_no_beam_search / _generate_beam_search loops that PretrainedBartModel inherits
                                          (src/model/mixins.py:10, :336-382; src/model/utils.py:1)
  modeling_utils: WEIGHTS_NAME & friends, cached_path / hf_bucket_url / is_remote_url (local
    paths only), PretrainedConfig (the installed transformers' class, so isinstance() checks on
    MultiModalBartConfig(BartConfig) hold)           (src/model/mixins.py:15-23)
  transformers.AdamW (HF-3.0.2 optimization.AdamW) (vcg_train.py:13, pretrain.py:13)
This file is an independent nn.Module-style restatement; oracle/kmbart_oracle.py is the
functional one.  Agreement of the two through the reference's own glue code is what
tests/test_oracle.py asserts.
"""
import logging
import math
import os
import random
import sys
import types
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

logger = logging.getLogger("transformers.generation_utils")

# HF-3.0.2 PretrainedConfig generation / output defaults that newer transformers moved elsewhere
HF302_CONFIG_DEFAULTS = dict(
    max_length=20, min_length=0, do_sample=False, early_stopping=False, num_beams=1, temperature=1.0,
    top_k=50, top_p=1.0, repetition_penalty=1.0, length_penalty=1.0, no_repeat_ngram_size=0,
    bad_words_ids=None, num_return_sequences=1, use_cache=True, output_attentions=False,
    output_hidden_states=False,
)


def apply_config_defaults(config):
    for k, v in HF302_CONFIG_DEFAULTS.items():
        if getattr(config, k, None) is None and not (k == "bad_words_ids"):
            setattr(config, k, v)
    if not hasattr(config, "bad_words_ids"):
        config.bad_words_ids = None
    return config


# ------------------------------------------------------------------ modeling_bart helpers
def invert_mask(attention_mask):
    assert attention_mask.dim() == 2
    return attention_mask.eq(0)


def shift_tokens_right(input_ids, pad_token_id):
    prev_output_tokens = input_ids.clone()
    index_of_eos = (input_ids.ne(pad_token_id).sum(dim=1) - 1).unsqueeze(-1)
    prev_output_tokens[:, 0] = input_ids.gather(1, index_of_eos).squeeze()
    prev_output_tokens[:, 1:] = input_ids[:, :-1]
    return prev_output_tokens


def make_padding_mask(input_ids, padding_idx=1):
    padding_mask = input_ids.eq(padding_idx)
    if not padding_mask.any():
        padding_mask = None
    return padding_mask


def fill_with_neg_inf(t):
    return t.float().fill_(float("-inf")).type_as(t)


def _prepare_bart_decoder_inputs(config, input_ids, decoder_input_ids=None, decoder_padding_mask=None,
                                 causal_mask_dtype=torch.float32):
    pad_token_id = config.pad_token_id
    if decoder_input_ids is None:
        decoder_input_ids = shift_tokens_right(input_ids, pad_token_id)
    bsz, tgt_len = decoder_input_ids.size()
    if decoder_padding_mask is None:
        decoder_padding_mask = make_padding_mask(decoder_input_ids, pad_token_id)
    else:
        decoder_padding_mask = invert_mask(decoder_padding_mask)
    causal_mask = torch.triu(fill_with_neg_inf(torch.zeros(tgt_len, tgt_len)), 1).to(
        dtype=causal_mask_dtype, device=decoder_input_ids.device)
    return decoder_input_ids, decoder_padding_mask, causal_mask


def _make_linear_from_emb(emb):
    vocab_size, emb_size = emb.weight.shape
    lin_layer = nn.Linear(vocab_size, emb_size, bias=False)
    lin_layer.weight.data = emb.weight.data
    return lin_layer


def _filter_out_falsey_values(tup) -> Tuple:
    return tuple(x for x in tup if isinstance(x, torch.Tensor) or x)


def _reorder_buffer(attn_cache, new_order):
    for k, input_buffer_k in attn_cache.items():
        if input_buffer_k is not None:
            attn_cache[k] = input_buffer_k.index_select(0, new_order)
    return attn_cache


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True):
    return torch.nn.LayerNorm(normalized_shape, eps, elementwise_affine)


class LearnedPositionalEmbedding(nn.Embedding):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: int, offset):
        self.offset = offset
        assert padding_idx is not None
        num_embeddings += offset
        super().__init__(num_embeddings, embedding_dim, padding_idx=padding_idx)

    def forward(self, input_ids, use_cache=False):
        bsz, seq_len = input_ids.shape[:2]
        if use_cache:
            positions = input_ids.data.new(1, 1).fill_(seq_len - 1)
        else:
            positions = torch.arange(seq_len, dtype=torch.long, device=self.weight.device)
        return super().forward(positions + self.offset)


class SinusoidalPositionalEmbedding(nn.Embedding):
    """Only referenced by the reference when config.static_position_embeddings (never on this path)."""

    def __init__(self, num_positions, embedding_dim, padding_idx=None):
        super().__init__(num_positions, embedding_dim)
        raise NotImplementedError("static position embeddings are outside the KM-BART configs")


class SelfAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, encoder_decoder_attention=False):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.encoder_decoder_attention = encoder_decoder_attention
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.cache_key = "encoder_decoder" if self.encoder_decoder_attention else "self"

    def _shape(self, tensor, dim_0, bsz):
        return tensor.contiguous().view(dim_0, bsz * self.num_heads, self.head_dim).transpose(0, 1)

    def forward(self, query, key: Optional[Tensor], key_padding_mask: Optional[Tensor] = None,
                layer_state: Optional[Dict[str, Optional[Tensor]]] = None, attn_mask: Optional[Tensor] = None,
                output_attentions=False):
        static_kv: bool = self.encoder_decoder_attention
        tgt_len, bsz, embed_dim = query.size()
        assert embed_dim == self.embed_dim
        if layer_state is not None:
            saved_state = layer_state.get(self.cache_key, {})
            if "prev_key" in saved_state and static_kv:
                key = None
        else:
            saved_state = None
            layer_state = {}
        q = self.q_proj(query) * self.scaling
        if static_kv:
            if key is None:
                k = v = None
            else:
                k = self.k_proj(key)
                v = self.v_proj(key)
        else:
            k = self.k_proj(query)
            v = self.v_proj(query)
        q = self._shape(q, tgt_len, bsz)
        if k is not None:
            k = self._shape(k, -1, bsz)
        if v is not None:
            v = self._shape(v, -1, bsz)
        if saved_state is not None:
            k, v, key_padding_mask = self._use_saved_state(k, v, saved_state, key_padding_mask, static_kv, bsz)
        layer_state[self.cache_key] = {
            "prev_key": k.view(bsz, self.num_heads, -1, self.head_dim),
            "prev_value": v.view(bsz, self.num_heads, -1, self.head_dim),
            "prev_key_padding_mask": key_padding_mask if not static_kv else None,
        }
        assert k is not None
        src_len = k.size(1)
        attn_weights = torch.bmm(q, k.transpose(1, 2))
        assert attn_weights.size() == (bsz * self.num_heads, tgt_len, src_len)
        if attn_mask is not None:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len) + attn_mask
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        if key_padding_mask is not None and key_padding_mask.dim() == 0:
            key_padding_mask = None
        assert key_padding_mask is None or key_padding_mask.size()[:2] == (bsz, src_len)
        if key_padding_mask is not None:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len)
            reshaped = key_padding_mask.unsqueeze(1).unsqueeze(2)
            attn_weights = attn_weights.masked_fill(reshaped, float("-inf"))
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        attn_weights = F.softmax(attn_weights, dim=-1)
        attn_probs = F.dropout(attn_weights, p=self.dropout, training=self.training)
        assert v is not None
        attn_output = torch.bmm(attn_probs, v)
        attn_output = attn_output.transpose(0, 1).contiguous().view(tgt_len, bsz, embed_dim)
        attn_output = self.out_proj(attn_output)
        if output_attentions:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len)
        else:
            attn_weights = None
        return attn_output, attn_weights

    def _use_saved_state(self, k, v, saved_state, key_padding_mask, static_kv, bsz):
        if "prev_key" in saved_state:
            prev_key = saved_state["prev_key"].view(bsz * self.num_heads, -1, self.head_dim)
            k = prev_key if static_kv else torch.cat([prev_key, k], dim=1)
        if "prev_value" in saved_state:
            prev_value = saved_state["prev_value"].view(bsz * self.num_heads, -1, self.head_dim)
            v = prev_value if static_kv else torch.cat([prev_value, v], dim=1)
        assert k is not None and v is not None
        prev_key_padding_mask = saved_state.get("prev_key_padding_mask", None)
        key_padding_mask = self._cat_prev_key_padding_mask(key_padding_mask, prev_key_padding_mask, bsz, k.size(1), static_kv)
        return k, v, key_padding_mask

    @staticmethod
    def _cat_prev_key_padding_mask(key_padding_mask, prev_key_padding_mask, batch_size, src_len, static_kv):
        if prev_key_padding_mask is not None:
            if static_kv:
                new_key_padding_mask = prev_key_padding_mask
            else:
                new_key_padding_mask = torch.cat([prev_key_padding_mask, key_padding_mask], dim=1)
        elif key_padding_mask is not None:
            filler = torch.zeros(batch_size, src_len - key_padding_mask.size(1), dtype=key_padding_mask.dtype,
                                 device=key_padding_mask.device)
            new_key_padding_mask = torch.cat([filler, key_padding_mask], dim=1)
        else:
            new_key_padding_mask = prev_key_padding_mask
        return new_key_padding_mask


def _act(name):
    assert name == "gelu", "KM-BART configs use activation_function='gelu' (exact erf GELU)"
    return F.gelu


class EncoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = SelfAttention(self.embed_dim, config.encoder_attention_heads, dropout=config.attention_dropout)
        self.normalize_before = config.normalize_before
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.dropout = config.dropout
        self.activation_fn = _act(config.activation_function)
        self.activation_dropout = config.activation_dropout
        self.fc1 = nn.Linear(self.embed_dim, config.encoder_ffn_dim)
        self.fc2 = nn.Linear(config.encoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)

    def forward(self, x, encoder_padding_mask, output_attentions=False):
        residual = x
        if self.normalize_before:
            x = self.self_attn_layer_norm(x)
        x, attn_weights = self.self_attn(query=x, key=x, key_padding_mask=encoder_padding_mask,
                                         output_attentions=output_attentions)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if not self.normalize_before:
            x = self.self_attn_layer_norm(x)
        residual = x
        if self.normalize_before:
            x = self.final_layer_norm(x)
        x = self.activation_fn(self.fc1(x))
        x = F.dropout(x, p=self.activation_dropout, training=self.training)
        x = self.fc2(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if not self.normalize_before:
            x = self.final_layer_norm(x)
        return x, attn_weights


class DecoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = SelfAttention(embed_dim=self.embed_dim, num_heads=config.decoder_attention_heads,
                                       dropout=config.attention_dropout)
        self.dropout = config.dropout
        self.activation_fn = _act(config.activation_function)
        self.activation_dropout = config.activation_dropout
        self.normalize_before = config.normalize_before
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.encoder_attn = SelfAttention(self.embed_dim, config.decoder_attention_heads, dropout=config.attention_dropout,
                                          encoder_decoder_attention=True)
        self.encoder_attn_layer_norm = LayerNorm(self.embed_dim)
        self.fc1 = nn.Linear(self.embed_dim, config.decoder_ffn_dim)
        self.fc2 = nn.Linear(config.decoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)

    def forward(self, x, encoder_hidden_states, encoder_attn_mask=None, layer_state=None, causal_mask=None,
                decoder_padding_mask=None, output_attentions=False):
        residual = x
        if layer_state is None:
            layer_state = {}
        if self.normalize_before:
            x = self.self_attn_layer_norm(x)
        x, self_attn_weights = self.self_attn(query=x, key=x, layer_state=layer_state,
                                              key_padding_mask=decoder_padding_mask, attn_mask=causal_mask,
                                              output_attentions=output_attentions)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if not self.normalize_before:
            x = self.self_attn_layer_norm(x)
        residual = x
        assert self.encoder_attn.cache_key != self.self_attn.cache_key
        if self.normalize_before:
            x = self.encoder_attn_layer_norm(x)
        x, _ = self.encoder_attn(query=x, key=encoder_hidden_states, key_padding_mask=encoder_attn_mask,
                                 layer_state=layer_state)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if not self.normalize_before:
            x = self.encoder_attn_layer_norm(x)
        residual = x
        if self.normalize_before:
            x = self.final_layer_norm(x)
        x = self.activation_fn(self.fc1(x))
        x = F.dropout(x, p=self.activation_dropout, training=self.training)
        x = self.fc2(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if not self.normalize_before:
            x = self.final_layer_norm(x)
        return x, self_attn_weights, layer_state


class BartDecoder(nn.Module):
    def __init__(self, config, embed_tokens: nn.Embedding):
        super().__init__()
        self.dropout = config.dropout
        self.layerdrop = config.decoder_layerdrop
        self.padding_idx = embed_tokens.padding_idx
        self.max_target_positions = config.max_position_embeddings
        self.embed_scale = math.sqrt(config.d_model) if config.scale_embedding else 1.0
        self.embed_tokens = embed_tokens
        if config.static_position_embeddings:
            self.embed_positions = SinusoidalPositionalEmbedding(config.max_position_embeddings, config.d_model, config.pad_token_id)
        else:
            self.embed_positions = LearnedPositionalEmbedding(config.max_position_embeddings, config.d_model,
                                                              self.padding_idx, config.extra_pos_embeddings)
        self.layers = nn.ModuleList([DecoderLayer(config) for _ in range(config.decoder_layers)])
        self.layernorm_embedding = LayerNorm(config.d_model) if config.normalize_embedding else nn.Identity()
        self.layer_norm = LayerNorm(config.d_model) if config.add_final_layer_norm else None

    def forward(self, input_ids, encoder_hidden_states, encoder_padding_mask, decoder_padding_mask, decoder_causal_mask,
                decoder_cached_states=None, use_cache=False, output_attentions=False, output_hidden_states=False, **unused):
        if encoder_padding_mask is not None:
            encoder_padding_mask = invert_mask(encoder_padding_mask)
        positions = self.embed_positions(input_ids, use_cache=use_cache)
        if use_cache:
            input_ids = input_ids[:, -1:]
            positions = positions[:, -1:]
        x = self.embed_tokens(input_ids) * self.embed_scale
        x += positions
        x = self.layernorm_embedding(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = x.transpose(0, 1)
        encoder_hidden_states = encoder_hidden_states.transpose(0, 1)
        all_hidden_states = ()
        all_self_attns = ()
        next_decoder_cache = []
        for idx, decoder_layer in enumerate(self.layers):
            if output_hidden_states:
                all_hidden_states += (x,)
            dropout_probability = random.uniform(0, 1)
            if self.training and (dropout_probability < self.layerdrop):
                continue
            layer_state = decoder_cached_states[idx] if decoder_cached_states is not None else None
            x, layer_self_attn, layer_past = decoder_layer(
                x, encoder_hidden_states, encoder_attn_mask=encoder_padding_mask, decoder_padding_mask=decoder_padding_mask,
                layer_state=layer_state, causal_mask=decoder_causal_mask, output_attentions=output_attentions)
            if use_cache:
                next_decoder_cache.append(layer_past.copy())
            if self.layer_norm and (idx == len(self.layers) - 1):
                x = self.layer_norm(x)
            if output_attentions:
                all_self_attns += (layer_self_attn,)
        all_hidden_states = [hidden_state.transpose(0, 1) for hidden_state in all_hidden_states]
        x = x.transpose(0, 1)
        encoder_hidden_states = encoder_hidden_states.transpose(0, 1)
        if use_cache:
            next_cache = ((encoder_hidden_states, encoder_padding_mask), next_decoder_cache)
        else:
            next_cache = None
        return x, next_cache, all_hidden_states, list(all_self_attns)


class BartClassificationHead(nn.Module):
    def __init__(self, input_dim, inner_dim, num_classes, pooler_dropout):
        super().__init__()
        self.dense = nn.Linear(input_dim, inner_dim)
        self.dropout = nn.Dropout(p=pooler_dropout)
        self.out_proj = nn.Linear(inner_dim, num_classes)

    def forward(self, x):
        x = self.dropout(x)
        x = self.dense(x)
        x = torch.tanh(x)
        x = self.dropout(x)
        x = self.out_proj(x)
        return x


# ------------------------------------------------------------------ generation_utils
def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("Inf"), min_tokens_to_keep=1):
    if top_k > 0:
        top_k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        indices_to_remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
        logits[indices_to_remove] = filter_value
    if top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        sorted_indices_to_remove = cumulative_probs > top_p
        if min_tokens_to_keep > 1:
            sorted_indices_to_remove[..., :min_tokens_to_keep] = 0
        sorted_indices_to_remove[..., 1:] = sorted_indices_to_remove[..., :-1].clone()
        sorted_indices_to_remove[..., 0] = 0
        indices_to_remove = sorted_indices_to_remove.scatter(1, sorted_indices, sorted_indices_to_remove)
        logits[indices_to_remove] = filter_value
    return logits


class BeamHypotheses(object):
    def __init__(self, num_beams, max_length, length_penalty, early_stopping):
        self.max_length = max_length - 1
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                sorted_scores = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[sorted_scores[0][1]]
                self.worst_score = sorted_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        elif self.early_stopping:
            return True
        else:
            cur_score = best_sum_logprobs / cur_len ** self.length_penalty
            return self.worst_score >= cur_score


class _GenerationLoops:
    """The parts of HF-3.0.2 generation_utils.GenerationMixin that the reference's classes inherit
    through PretrainedBartModel and call from src/model/mixins.py:336-382."""

    def _use_cache(self, outputs, use_cache):
        if len(outputs) <= 1 or use_cache is False:
            return False
        if hasattr(self.config, "mem_len") and self.config.mem_len == 0:
            return False
        return True

    def enforce_repetition_penalty_(self, lprobs, batch_size, num_beams, prev_output_tokens, repetition_penalty):
        for i in range(batch_size * num_beams):
            for previous_token in set(prev_output_tokens[i].tolist()):
                if lprobs[i, previous_token] < 0:
                    lprobs[i, previous_token] *= repetition_penalty
                else:
                    lprobs[i, previous_token] /= repetition_penalty

    def postprocess_next_token_scores(self, scores, input_ids, no_repeat_ngram_size, bad_words_ids, cur_len, min_length,
                                      max_length, eos_token_id, repetition_penalty, batch_size, num_beams):
        if repetition_penalty != 1.0:
            self.enforce_repetition_penalty_(scores, batch_size, num_beams, input_ids, repetition_penalty)
        if eos_token_id is not None and cur_len < min_length:
            scores[:, eos_token_id] = -float("inf")
        if no_repeat_ngram_size > 0 or bad_words_ids is not None:
            raise NotImplementedError("n-gram / bad-word bans are not reached by the reference's call sites")
        return scores

    def _generate_no_beam_search(self, input_ids, cur_len, max_length, min_length, do_sample, temperature, top_k, top_p,
                                 repetition_penalty, no_repeat_ngram_size, bad_words_ids, pad_token_id, eos_token_id,
                                 batch_size, encoder_outputs, attention_mask, use_cache, model_specific_kwargs):
        unfinished_sents = input_ids.new(batch_size).fill_(1)
        sent_lengths = input_ids.new(batch_size).fill_(max_length)
        past = (encoder_outputs, None) if encoder_outputs is not None else None
        while cur_len < max_length:
            model_inputs = self.prepare_inputs_for_generation(input_ids, past=past, attention_mask=attention_mask,
                                                              use_cache=use_cache, **model_specific_kwargs)
            outputs = self(**model_inputs)
            next_token_logits = outputs[0][:, -1, :]
            scores = self.postprocess_next_token_scores(
                scores=next_token_logits, input_ids=input_ids, no_repeat_ngram_size=no_repeat_ngram_size,
                bad_words_ids=bad_words_ids, cur_len=cur_len, min_length=min_length, max_length=max_length,
                eos_token_id=eos_token_id, repetition_penalty=repetition_penalty, batch_size=batch_size, num_beams=1)
            if self._use_cache(outputs, use_cache):
                past = outputs[1]
            if do_sample:
                if temperature != 1.0:
                    scores = scores / temperature
                next_token_logscores = top_k_top_p_filtering(scores, top_k=top_k, top_p=top_p)
                probs = F.softmax(next_token_logscores, dim=-1)
                next_token = torch.multinomial(probs, num_samples=1).squeeze(1)
            else:
                next_token = torch.argmax(next_token_logits, dim=-1)
            if eos_token_id is not None:
                tokens_to_add = next_token * unfinished_sents + (pad_token_id) * (1 - unfinished_sents)
            else:
                tokens_to_add = next_token
            input_ids = torch.cat([input_ids, tokens_to_add.unsqueeze(-1)], dim=-1)
            cur_len = cur_len + 1
            if eos_token_id is not None:
                eos_in_sents = tokens_to_add == eos_token_id
                is_sents_unfinished_and_token_to_add_is_eos = unfinished_sents.mul(eos_in_sents.long()).bool()
                sent_lengths.masked_fill_(is_sents_unfinished_and_token_to_add_is_eos, cur_len)
                unfinished_sents.mul_((~eos_in_sents).long())
            if unfinished_sents.max() == 0:
                break
        if sent_lengths.min().item() != sent_lengths.max().item():
            assert pad_token_id is not None
            decoded = input_ids.new(batch_size, sent_lengths.max().item()).fill_(pad_token_id)
        else:
            decoded = input_ids
        for hypo_idx, hypo in enumerate(input_ids):
            decoded[hypo_idx, : sent_lengths[hypo_idx]] = hypo[: sent_lengths[hypo_idx]]
        return decoded

    def _generate_beam_search(self, input_ids, cur_len, max_length, min_length, do_sample, early_stopping, temperature,
                              top_k, top_p, repetition_penalty, no_repeat_ngram_size, bad_words_ids, pad_token_id,
                              eos_token_id, batch_size, num_return_sequences, length_penalty, num_beams, vocab_size,
                              encoder_outputs, attention_mask, use_cache, model_specific_kwargs):
        generated_hyps = [BeamHypotheses(num_beams, max_length, length_penalty, early_stopping=early_stopping)
                          for _ in range(batch_size)]
        beam_scores = torch.zeros((batch_size, num_beams), dtype=torch.float, device=input_ids.device)
        if do_sample is False:
            beam_scores[:, 1:] = -1e9
        beam_scores = beam_scores.view(-1)
        past = (encoder_outputs, None) if encoder_outputs is not None else None
        done = [False for _ in range(batch_size)]
        while cur_len < max_length:
            model_inputs = self.prepare_inputs_for_generation(input_ids, past=past, attention_mask=attention_mask,
                                                              use_cache=use_cache, **model_specific_kwargs)
            outputs = self(**model_inputs)
            next_token_logits = outputs[0][:, -1, :]
            if self._use_cache(outputs, use_cache):
                past = outputs[1]
            if self.config.is_encoder_decoder and do_sample is False:
                next_token_logits = self.adjust_logits_during_generation(next_token_logits, cur_len=cur_len,
                                                                         max_length=max_length)
            scores = F.log_softmax(next_token_logits, dim=-1)
            scores = self.postprocess_next_token_scores(
                scores=scores, input_ids=input_ids, no_repeat_ngram_size=no_repeat_ngram_size, bad_words_ids=bad_words_ids,
                cur_len=cur_len, min_length=min_length, max_length=max_length, eos_token_id=eos_token_id,
                repetition_penalty=repetition_penalty, batch_size=batch_size, num_beams=num_beams)
            assert scores.shape == (batch_size * num_beams, vocab_size)
            if do_sample:
                _scores = scores + beam_scores[:, None].expand_as(scores)
                if temperature != 1.0:
                    _scores = _scores / temperature
                _scores = top_k_top_p_filtering(_scores, top_k=top_k, top_p=top_p, min_tokens_to_keep=2)
                _scores = _scores.contiguous().view(batch_size, num_beams * vocab_size)
                probs = F.softmax(_scores, dim=-1)
                next_tokens = torch.multinomial(probs, num_samples=2 * num_beams)
                next_scores = torch.gather(_scores, -1, next_tokens)
                next_scores, next_scores_indices = torch.sort(next_scores, descending=True, dim=1)
                next_tokens = torch.gather(next_tokens, -1, next_scores_indices)
            else:
                next_scores = scores + beam_scores[:, None].expand_as(scores)
                next_scores = next_scores.view(batch_size, num_beams * vocab_size)
                next_scores, next_tokens = torch.topk(next_scores, 2 * num_beams, dim=1, largest=True, sorted=True)
            assert next_scores.size() == next_tokens.size() == (batch_size, 2 * num_beams)
            next_batch_beam = []
            for batch_idx in range(batch_size):
                if done[batch_idx]:
                    assert len(generated_hyps[batch_idx]) >= num_beams
                    assert eos_token_id is not None and pad_token_id is not None
                    next_batch_beam.extend([(0, pad_token_id, 0)] * num_beams)
                    continue
                next_sent_beam = []
                for beam_token_rank, (beam_token_id, beam_token_score) in enumerate(
                        zip(next_tokens[batch_idx], next_scores[batch_idx])):
                    beam_id = beam_token_id // vocab_size
                    token_id = beam_token_id % vocab_size
                    effective_beam_id = batch_idx * num_beams + beam_id
                    if (eos_token_id is not None) and (token_id.item() == eos_token_id):
                        is_beam_token_worse_than_top_num_beams = beam_token_rank >= num_beams
                        if is_beam_token_worse_than_top_num_beams:
                            continue
                        generated_hyps[batch_idx].add(input_ids[effective_beam_id].clone(), beam_token_score.item())
                    else:
                        next_sent_beam.append((beam_token_score, token_id, effective_beam_id))
                    if len(next_sent_beam) == num_beams:
                        break
                done[batch_idx] = done[batch_idx] or generated_hyps[batch_idx].is_done(
                    next_scores[batch_idx].max().item(), cur_len=cur_len)
                assert len(next_sent_beam) == num_beams, "Beam should always be full"
                next_batch_beam.extend(next_sent_beam)
                assert len(next_batch_beam) == num_beams * (batch_idx + 1)
            if all(done):
                break
            assert len(next_batch_beam) == batch_size * num_beams
            beam_scores = beam_scores.new([x[0] for x in next_batch_beam])
            beam_tokens = input_ids.new([x[1] for x in next_batch_beam])
            beam_idx = input_ids.new([x[2] for x in next_batch_beam])
            input_ids = input_ids[beam_idx, :]
            input_ids = torch.cat([input_ids, beam_tokens.unsqueeze(1)], dim=-1)
            cur_len = cur_len + 1
            if past is not None:
                past = self._reorder_cache(past, beam_idx)
        for batch_idx in range(batch_size):
            if done[batch_idx]:
                continue
            if eos_token_id is not None and all(
                    (token_id % vocab_size).item() != eos_token_id for token_id in next_tokens[batch_idx]):
                assert torch.all(next_scores[batch_idx, :num_beams] == beam_scores.view(batch_size, num_beams)[batch_idx])
            for beam_id in range(num_beams):
                effective_beam_id = batch_idx * num_beams + beam_id
                final_score = beam_scores[effective_beam_id].item()
                final_tokens = input_ids[effective_beam_id]
                generated_hyps[batch_idx].add(final_tokens, final_score)
        output_batch_size = batch_size if do_sample else batch_size * num_return_sequences
        output_num_return_sequences_per_batch = 1 if do_sample else num_return_sequences
        sent_lengths = input_ids.new(output_batch_size)
        best = []
        for i, hypotheses in enumerate(generated_hyps):
            sorted_hyps = sorted(hypotheses.beams, key=lambda x: x[0])
            for j in range(output_num_return_sequences_per_batch):
                effective_batch_idx = output_num_return_sequences_per_batch * i + j
                best_hyp = sorted_hyps.pop()[1]
                sent_lengths[effective_batch_idx] = len(best_hyp)
                best.append(best_hyp)
        if sent_lengths.min().item() != sent_lengths.max().item():
            assert pad_token_id is not None
            sent_max_len = min(sent_lengths.max().item() + 1, max_length)
            decoded = input_ids.new(output_batch_size, sent_max_len).fill_(pad_token_id)
            for i, hypo in enumerate(best):
                decoded[i, : sent_lengths[i]] = hypo
                if sent_lengths[i] < max_length:
                    decoded[i, sent_lengths[i]] = eos_token_id
        else:
            assert (len(hypo) == max_length for hypo in best)
            decoded = torch.stack(best).type(torch.long).to(next(self.parameters()).device)
        return decoded


# ------------------------------------------------------------------ PretrainedBartModel
class PretrainedBartModel(_GenerationLoops, nn.Module):
    """HF-3.0.2 PreTrainedModel + PretrainedBartModel, reduced to what the reference touches."""
    base_model_prefix = "model"
    config_class = None  # set in install() to the installed transformers' BartConfig

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = apply_config_defaults(config)

    @property
    def base_model(self):
        return getattr(self, self.base_model_prefix, self)

    def _init_weights(self, module):
        std = self.config.init_std
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    def init_weights(self):
        self.apply(self._init_weights)
        self.tie_weights()

    def tie_weights(self):
        # HF-3.0.2 ties get_output_embeddings() to get_input_embeddings(); for BART the output
        # embedding is made on the fly from `shared`, so there is nothing to re-point.
        pass

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dummy_inputs(self):
        pad_token = self.config.pad_token_id
        input_ids = torch.tensor([[0, 6, 10, 4, 2], [0, 8, 12, 2, pad_token]], device=self.device)
        return {"attention_mask": input_ids.ne(pad_token), "input_ids": input_ids}


# ------------------------------------------------------------------ optimization.AdamW
class AdamW(torch.optim.Optimizer):
    """HF-3.0.2 transformers.AdamW (what vcg_train.py:13,100 imports)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)

    def step(self, closure=None):
        loss = None
        if closure is not None:
            loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.data
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p.data)
                    state["exp_avg_sq"] = torch.zeros_like(p.data)
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                beta1, beta2 = group["betas"]
                state["step"] += 1
                exp_avg.mul_(beta1).add_(grad, alpha=1.0 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
                denom = exp_avg_sq.sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    bias_correction1 = 1.0 - beta1 ** state["step"]
                    bias_correction2 = 1.0 - beta2 ** state["step"]
                    step_size = step_size * math.sqrt(bias_correction2) / bias_correction1
                p.data.addcdiv_(exp_avg, denom, value=-step_size)
                if group["weight_decay"] > 0.0:
                    p.data.add_(p.data, alpha=-group["lr"] * group["weight_decay"])
        return loss


# ------------------------------------------------------------------ modeling_utils
def is_remote_url(url_or_filename):
    return str(url_or_filename).startswith(("http://", "https://", "s3://"))


def hf_bucket_url(model_id, filename, use_cdn=True):
    raise EnvironmentError("no network: only local checkpoints can be loaded ({})".format(model_id))


def cached_path(url_or_filename, cache_dir=None, force_download=False, proxies=None, resume_download=False,
                user_agent=None, local_files_only=False, **kw):
    if os.path.exists(url_or_filename):
        return url_or_filename
    raise EnvironmentError("file {} not found".format(url_or_filename))


def install(reference_root="/root/reference"):
    """Register the stand-in modules.  Idempotent."""
    import transformers
    if "transformers.modeling_bart" not in sys.modules:
        me = sys.modules[__name__]
        mb = types.ModuleType("transformers.modeling_bart")
        for n in ("PretrainedBartModel", "BartDecoder", "DecoderLayer", "EncoderLayer", "SelfAttention",
                  "LearnedPositionalEmbedding", "SinusoidalPositionalEmbedding", "LayerNorm", "BartClassificationHead",
                  "invert_mask", "_prepare_bart_decoder_inputs", "_make_linear_from_emb", "_filter_out_falsey_values",
                  "_reorder_buffer", "shift_tokens_right", "make_padding_mask"):
            setattr(mb, n, getattr(me, n))
        gu = types.ModuleType("transformers.generation_utils")
        gu.logger, gu.Iterable = logger, Iterable
        gu.top_k_top_p_filtering, gu.BeamHypotheses = top_k_top_p_filtering, BeamHypotheses
        mu = types.ModuleType("transformers.modeling_utils_hf302")
        real_mu = sys.modules.get("transformers.modeling_utils")
        try:
            from transformers import PretrainedConfig
        except Exception:  # pragma: no cover
            from transformers import PreTrainedConfig as PretrainedConfig
        names = dict(hf_bucket_url=hf_bucket_url, cached_path=cached_path, TF2_WEIGHTS_NAME="tf_model.h5",
                     WEIGHTS_NAME="pytorch_model.bin", TF_WEIGHTS_NAME="model.ckpt", is_remote_url=is_remote_url,
                     PretrainedConfig=PretrainedConfig)
        # keep the real modeling_utils usable: graft the missing legacy names onto it
        import transformers.modeling_utils as real_mu  # noqa: F811
        for k, v in names.items():
            if not hasattr(real_mu, k):
                setattr(real_mu, k, v)
        PretrainedBartModel.config_class = transformers.BartConfig
        sys.modules["transformers.modeling_bart"] = mb
        sys.modules["transformers.generation_utils"] = gu
        transformers.modeling_bart = mb
        transformers.generation_utils = gu
        if not hasattr(transformers, "AdamW"):
            try:
                transformers.AdamW = AdamW
            except Exception:  # lazy-module attribute guard
                pass
    # sys.path is NOT touched here: import_reference() puts the checkout first only for the duration of its imports,
    # so the product's same-named `src` package keeps resolving everywhere else (incl. spawned test workers)
    return os.path.isdir(reference_root) if reference_root else False


def import_reference(reference_root="/root/reference"):
    """Import the reference's own src.model package (NOT the product's same-named drop-in).
    Returns the modules (config, model, modules, mixins) or raises if the checkout is absent."""
    if not os.path.isdir(reference_root):
        raise FileNotFoundError(reference_root)
    install(reference_root)
    import importlib
    # the product package is also called `src`; make sure the reference's wins for this import
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    saved_path = list(sys.path)
    try:
        sys.path[:] = [reference_root] + [p for p in sys.path if not p.rstrip("/").endswith("km-bart_b200")]
        mods = {n: importlib.import_module("src.model." + n) for n in ("config", "modules", "mixins", "model", "utils")}
    finally:
        ref_loaded = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
        sys.modules.update(saved)
        sys.path[:] = saved_path
    mods["_loaded"] = ref_loaded
    return mods
