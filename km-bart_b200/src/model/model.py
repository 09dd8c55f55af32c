"""KM-BART model classes — drop-in for the reference's src/model/model.py.

Same classes, constructor/forward signatures, return tuples, state_dict keys and parameters()
order as MultiModalBartModel (:27-114), MultiModalBartForPreTraining (:125-309),
MultiModalBartForConditionalGeneration (:317-405) and ReasoningClassification (:408-426).
The arithmetic runs in hand-written sm_100a kernels through kmbart.engine.Engine:
  * labels given  -> one fused forward (+ stashed activations) whose loss node backpropagates
    through a single autograd.Function that launches the whole backward plan; the LM-head logits
    are never written to HBM (outputs[1] is a LazyLogits that materialises only if touched);
  * no labels, use_cache=False -> full-sequence inference plan, logits materialised;
  * use_cache=True -> one KV-cached decode step returning the legacy cache structure
    ((enc_out, enc_padding_mask), [ {self:{prev_key,...}, encoder_decoder:{...}} ... ]).
There is no PyTorch fallback: on a non-sm_100 device the engine constructor raises."""
import weakref

import torch
import torch.nn.functional as F
from torch import nn

from src.model.config import MultiModalBartConfig
from src.model.mixins import GenerationMixin, FromPretrainedMixin
from src.model.modules import MultiModalBartEncoder, BartDecoder, BartClassificationHead


class LazyLogits:
    """outputs[1] of a labelled forward: [B, T, V] logits computed on first use from a private copy
    of the decoder states (the reference materialises 1.24 GB of fp32 logits every step,
    src/model/model.py:397; its callers touch them once per 100 steps, pretrain.py:285)."""

    def __init__(self, fn, shape, device):
        self._fn, self._shape, self._device, self._value = fn, torch.Size(shape), device, None

    def materialize(self):
        if self._value is None:
            self._value = self._fn()
            self._fn = None
        return self._value

    shape = property(lambda self: self._shape)
    device = property(lambda self: self._device)
    dtype = property(lambda self: torch.float32)

    def size(self, dim=None):
        return self._shape if dim is None else self._shape[dim]

    def dim(self):
        return len(self._shape)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda x: x.materialize() if isinstance(x, LazyLogits) else x
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)


class _FusedTrainStep(torch.autograd.Function):
    """loss = f(params); backward launches the engine's backward plan and hands each parameter its
    slice of the flat gradient buffer (so DDP / GradScaler / any optimizer see ordinary .grad)."""

    @staticmethod
    def forward(ctx, owner, arena, key, *params):
        ctx.owner, ctx.arena, ctx.key, ctx.serial = owner, arena, key, arena["serial"]
        ctx.n_params = len(params)
        return arena["loss"].clone().squeeze(0)

    @staticmethod
    def backward(ctx, grad_loss):
        owner = ctx.owner
        eng = owner._engine()
        # decided here, not at forward time: the reference's loop calls optimizer.zero_grad() between
        # forward and backward (src/training.py:134-142)
        accumulate = eng.grads_alias_flat_buffer()
        eng.train_backward(ctx.arena, ctx.key, grad_loss, accumulate, serial=ctx.serial)
        store = eng.store
        if accumulate:
            # gradients were added in place into the buffer the existing .grad tensors alias
            return (None, None, None) + (None,) * ctx.n_params
        grads = []
        for name, p in owner._named_params_cache:
            grads.append(store.grad_view(name) if p.requires_grad else None)
        return (None, None, None) + tuple(grads)


class PretrainedBartModel(nn.Module):
    """The slice of HF-3.0.2 PreTrainedModel/PretrainedBartModel the reference relies on."""
    config_class = MultiModalBartConfig
    base_model_prefix = "model"

    def __init__(self, config):
        super().__init__()
        self.config = config
        self._eng = None
        self._extra_backward = None

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

    def train(self, mode=True):
        """The fused kernels implement `dropout` (embeddings, residual branches) but neither dropout on the attention
        probabilities nor on the FFN activation (vcg_train.py:78-83 exposes both as flags): refuse to TRAIN with them
        here, when the script switches to training mode, instead of computing a different model."""
        cfg = getattr(self, "config", None)
        if mode and cfg is not None:
            for knob in ("attention_dropout", "activation_dropout"):
                if float(getattr(cfg, knob, 0.0) or 0.0) > 0.0:
                    raise ValueError(f"config.{knob} = {getattr(cfg, knob)} is not implemented by the B200 kernels (only `dropout` is); "
                                     f"set it to 0 for training — inference ignores it")
        return super().train(mode)

    def tie_weights(self):
        pass  # embeddings are tied by construction (one nn.Embedding object shared by encoder and decoder)

    @property
    def base_model(self):
        return getattr(self, self.base_model_prefix, self)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dummy_inputs(self):
        pad = self.config.pad_token_id
        ids = torch.tensor([[0, 6, 10, 4, 2], [0, 8, 12, 2, pad]], device=self.device)
        return {"attention_mask": ids.ne(pad), "input_ids": ids}

    _engine_prefix = "model."

    @property
    def precision(self):
        """'bf16' (default: tensor-core fast path) or 'fp32' (3xTF32 parity mode, inference only; kmbart/fp32.py)."""
        import os
        return os.environ.get("KMBART_PRECISION") or getattr(self.config, "kmb_precision", "bf16")

    def _fp32(self):
        eng = self._engine()
        if getattr(eng, "_fp32_path", None) is None:
            from kmbart.fp32 import Fp32Path
            eng._fp32_path = Fp32Path(eng)
        return eng._fp32_path

    def _engine(self):
        from kmbart.engine import Engine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("KM-BART (B200-native) runs on an sm_100 CUDA device only; there is no CPU path "
                               "(call .to('cuda') first)")
        if self._eng is None or self._eng.device != dev:
            self._eng = Engine(self, self.config, prefix=self._engine_prefix)
            self._named_params_cache = list(self.named_parameters())
        return self._eng


def _shift_tokens_right(input_ids, pad_token_id):
    """HF-3.0.2 shift_tokens_right: used when decoder_input_ids is None (src/model/model.py:63-70)."""
    prev = input_ids.clone()
    last = (input_ids.ne(pad_token_id).sum(dim=1) - 1).unsqueeze(-1)
    prev[:, 0] = input_ids.gather(1, last).squeeze(-1)
    prev[:, 1:] = input_ids[:, :-1]
    return prev


def _core_forward(owner, input_ids, image_features, attention_mask, decoder_input_ids, encoder_outputs,
                  decoder_attention_mask, decoder_cached_states, use_cache, output_attentions, output_hidden_states):
    """Shared inference body of MultiModalBartModel.forward (src/model/model.py:39-103).
    Returns (dec_hidden_f32 [B,T,d], dec_hidden_b16 [B*T,d], cache_or_None, enc_out [B,S,d])."""
    cfg = owner.config
    if decoder_input_ids is None:
        use_cache = False
    if output_attentions or output_hidden_states or cfg.output_attentions or cfg.output_hidden_states:
        raise NotImplementedError("attention maps / per-layer states are never materialised by the fused kernels")
    use_cache = use_cache if use_cache is not None else cfg.use_cache
    eng = owner._engine()
    if owner.precision == "fp32":
        return _core_forward_fp32(owner, input_ids, image_features, attention_mask, decoder_input_ids, encoder_outputs,
                                  decoder_attention_mask, decoder_cached_states, use_cache)
    if not use_cache:
        if decoder_input_ids is None:
            decoder_input_ids = _shift_tokens_right(input_ids, cfg.pad_token_id)
        if encoder_outputs is None:
            enc, dec, a = eng.infer_forward(input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask)
            return dec.clone(), a["dec_b16"], None, enc.clone()
        assert isinstance(encoder_outputs, tuple)
        dec, a = eng.decoder_full(encoder_outputs[0], attention_mask, decoder_input_ids, decoder_attention_mask)
        return dec.clone(), a["dec_b16"], None, encoder_outputs[0]
    # ---- cached single step
    if encoder_outputs is None:
        enc, _, _ = eng.infer_forward(input_ids, image_features, attention_mask, None, None, encoder_only=True)
        encoder_outputs = (enc.clone(),)
    assert isinstance(encoder_outputs, tuple)
    enc_out = encoder_outputs[0]
    enc_pad = attention_mask.eq(0) if attention_mask is not None else None
    pad_u8 = enc_pad.to(torch.uint8).contiguous() if enc_pad is not None else None
    T = decoder_input_ids.shape[1]
    h_b16, h_f32, caches = eng.decoder_step(decoder_input_ids[:, -1], T - 1, enc_out, pad_u8, decoder_cached_states)
    return h_f32.unsqueeze(1), h_b16, ((enc_out, enc_pad), caches), enc_out


def _core_forward_fp32(owner, input_ids, image_features, attention_mask, decoder_input_ids, encoder_outputs,
                       decoder_attention_mask, decoder_cached_states, use_cache):
    """fp32 parity mode of _core_forward: same control flow, kernels from kmbart/fp32.py (hidden states stay fp32)."""
    cfg = owner.config
    if hasattr(image_features, "as_list"):   # kmbart.feed.PackedImageFeatures
        image_features = image_features.as_list()
    fp = owner._fp32()
    if encoder_outputs is None:
        enc = fp.encoder(input_ids, image_features, attention_mask)
    else:
        assert isinstance(encoder_outputs, tuple)
        enc = encoder_outputs[0]
    if not use_cache:
        if decoder_input_ids is None:
            decoder_input_ids = _shift_tokens_right(input_ids, cfg.pad_token_id)
        dec = fp.decoder_full(enc, attention_mask, decoder_input_ids, decoder_attention_mask)
        return dec, dec.reshape(-1, dec.shape[-1]), None, enc
    enc_pad = attention_mask.eq(0) if attention_mask is not None else None
    pad_u8 = enc_pad.to(torch.uint8).contiguous() if enc_pad is not None else None
    T = decoder_input_ids.shape[1]
    h, caches = fp.decoder_step(decoder_input_ids[:, -1], T - 1, enc, pad_u8, decoder_cached_states)
    return h.unsqueeze(1), h, ((enc, enc_pad), caches), enc


class MultiModalBartModel(FromPretrainedMixin, PretrainedBartModel):
    _engine_prefix = ""

    def __init__(self, config: MultiModalBartConfig):
        super().__init__(config)
        padding_idx, vocab_size = config.pad_token_id, config.vocab_size
        self.shared = nn.Embedding(vocab_size, config.d_model, padding_idx)
        self.encoder = MultiModalBartEncoder(config, self.shared)
        self.decoder = BartDecoder(config, self.shared)
        self.encoder._owner_ref = weakref.ref(self)
        self._owner_ref = None
        self.init_weights()

    def _engine(self):
        owner = self._owner_ref() if self._owner_ref is not None else None
        return owner._engine() if owner is not None else super()._engine()

    def forward(self, input_ids, image_features, attention_mask=None, decoder_input_ids=None, encoder_outputs=None,
                decoder_attention_mask=None, decoder_cached_states=None, use_cache=None, output_attentions=None,
                output_hidden_states=None):
        if torch.is_grad_enabled() and self.training:
            raise NotImplementedError("train through MultiModalBartForConditionalGeneration / ForPreTraining "
                                      "(fused loss path); the bare model is inference-only")
        dec, _, cache, enc = _core_forward(self, input_ids, image_features, attention_mask, decoder_input_ids,
                                           encoder_outputs, decoder_attention_mask, decoder_cached_states, use_cache,
                                           output_attentions, output_hidden_states)
        out = (dec,) + ((cache,) if cache is not None else ()) + (enc,)
        return out

    def get_input_embeddings(self):
        return self.shared

    def set_input_embeddings(self, value):
        self.shared = value
        self.encoder.embed_tokens = self.shared
        self.decoder.embed_tokens = self.shared

    def get_output_embeddings(self):
        lin = nn.Linear(self.shared.weight.shape[1], self.shared.weight.shape[0], bias=False)
        lin.weight.data = self.shared.weight.data
        return lin


class _LMBase(FromPretrainedMixin, GenerationMixin, PretrainedBartModel):
    base_model_prefix = "model"

    def _finish_init(self):
        self.register_buffer("final_logits_bias", torch.zeros((1, self.model.shared.num_embeddings)))
        self.model.encoder._owner_ref = weakref.ref(self)
        self.model._owner_ref = weakref.ref(self)

    def _logits(self, h_b16, B, T):
        if self.precision == "fp32":
            return self._fp32().logits(h_b16, self.final_logits_bias).view(B, T, self.config.vocab_size)
        return self._engine().logits_from_hidden(h_b16, self.final_logits_bias).view(B, T, self.config.vocab_size)

    def _inference_outputs(self, input_ids, image_features, attention_mask, encoder_outputs, decoder_input_ids,
                           decoder_attention_mask, decoder_cached_states, use_cache, output_attentions, output_hidden_states):
        dec, h_b16, cache, enc = _core_forward(self, input_ids, image_features, attention_mask, decoder_input_ids,
                                               encoder_outputs, decoder_attention_mask, decoder_cached_states, use_cache,
                                               output_attentions, output_hidden_states)
        B, T = dec.shape[0], dec.shape[1]
        logits = self._logits(h_b16, B, T)
        return (logits,) + ((cache,) if cache is not None else ()) + (enc,), dec

    def _fused_loss(self, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels,
                    lm_factor):
        """Returns (lm_loss tensor with autograd node, LazyLogits, enc_out view, arena)."""
        cfg = self.config
        eng = self._engine()
        if decoder_input_ids is None:
            decoder_input_ids = _shift_tokens_right(input_ids, cfg.pad_token_id)
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        a = eng.train_forward(input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels,
                              self.final_logits_bias, self.training, lm_factor=lm_factor)
        key = eng.last_train[1]
        if grad:
            params = [p for _, p in self._named_params_cache]
            loss = _FusedTrainStep.apply(self, a, key, *params)
        else:
            loss = a["loss"].clone().squeeze(0)
        B, Sd, d = a["B"], a["Sd"], cfg.d_model
        h_copy = a["dec_b16"].clone()
        lazy = LazyLogits(lambda: self._logits(h_copy, B, Sd), (B, Sd, cfg.vocab_size), h_copy.device)
        enc = a["enc_f32"].view(B, a["Se"], d)
        return loss, lazy, enc, a


class MultiModalBartForConditionalGeneration(_LMBase):
    def __init__(self, config: MultiModalBartConfig):
        super().__init__(config)
        self.model = MultiModalBartModel(config)
        self._finish_init()

    def forward(self, input_ids, image_features, attention_mask=None, encoder_outputs=None, decoder_input_ids=None,
                decoder_attention_mask=None, decoder_cached_states=None, labels=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, **unused):
        if labels is not None:
            use_cache = False
            if self.precision == "fp32" and torch.is_grad_enabled() and self.training:
                raise NotImplementedError("fp32 parity mode is inference-only; train with the default bf16 path")
            if encoder_outputs is None and decoder_cached_states is None and self.precision != "fp32":
                loss, lazy, enc, _ = self._fused_loss(input_ids, image_features, attention_mask, decoder_input_ids,
                                                      decoder_attention_mask, labels, 1.0)
                return (loss, lazy, enc)
        outputs, _ = self._inference_outputs(input_ids, image_features, attention_mask, encoder_outputs, decoder_input_ids,
                                             decoder_attention_mask, decoder_cached_states, use_cache, output_attentions,
                                             output_hidden_states)
        if labels is not None:  # labelled call with pre-computed encoder states: evaluation-only loss
            lm_logits = outputs[0]
            loss = F.cross_entropy(lm_logits.view(-1, self.config.vocab_size), labels.view(-1))
            outputs = (loss,) + outputs
        return outputs


class MultiModalBartForPreTraining(_LMBase):
    def __init__(self, config: MultiModalBartConfig):
        super().__init__(config)
        self.cls_token_id = config.cls_token_id
        self.model = MultiModalBartModel(config)
        self.mrm_head = BartClassificationHead(config.d_model, config.d_model, config.num_labels, config.classif_dropout)
        self._init_weights(self.mrm_head.dense)
        self._init_weights(self.mrm_head.out_proj)
        self.attribute_head = BartClassificationHead(config.d_model, config.d_model, config.num_attributes, config.classif_dropout)
        self._init_weights(self.attribute_head.dense)
        self._init_weights(self.attribute_head.out_proj)
        self.relation_head = BartClassificationHead(config.d_model * 2, config.d_model, config.num_relations, config.classif_dropout)
        self._init_weights(self.relation_head.dense)
        self._init_weights(self.relation_head.out_proj)
        self._finish_init()

    def forward(self, input_ids, image_features, attention_mask=None, encoder_outputs=None, decoder_input_ids=None,
                decoder_attention_mask=None, decoder_cached_states=None, labels=None, mrm_labels=None, mrm_mask=None,
                attribute_labels=None, attribute_mask=None, relation_labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, **unused):
        any_label = (labels is not None) or (mrm_labels is not None) or (attribute_labels is not None) or \
                    (relation_labels is not None)
        if any_label:
            use_cache = False
        if (mrm_labels is not None) and (mrm_mask is None):
            raise ValueError('"mrm_mask" cannot be None while "mrm_labels" is set')
        if not any_label:
            outputs, _ = self._inference_outputs(input_ids, image_features, attention_mask, encoder_outputs,
                                                 decoder_input_ids, decoder_attention_mask, decoder_cached_states,
                                                 use_cache, output_attentions, output_hidden_states)
            return outputs
        from kmbart.heads import pretraining_forward
        return pretraining_forward(self, input_ids, image_features, attention_mask, decoder_input_ids,
                                   decoder_attention_mask, labels, mrm_labels, mrm_mask, attribute_labels, attribute_mask,
                                   relation_labels)


class ReasoningClassification(nn.Module):
    """Off the hot path (only scripts/prepare_atomic.py trains it); kept for API completeness
    (src/model/model.py:408-426) as a plain torch module."""

    def __init__(self, txt_dim, image_dim, inner_dim):
        super().__init__()
        self._txt_dim = txt_dim
        self._image_dim = image_dim
        self.txt_proj = nn.Linear(txt_dim, inner_dim)
        self.image_proj = nn.Linear(image_dim, inner_dim)
        self.out_proj = nn.Linear(2 * inner_dim, 2)
        self.act_fct = nn.Tanh()
        self.loss_fct = nn.CrossEntropyLoss()

    def forward(self, txt, image, label):
        txt_x = self.act_fct(self.txt_proj(txt.view(-1, self._txt_dim)))
        image_x = self.act_fct(self.image_proj(image.view(-1, self._image_dim)))
        x = self.out_proj(torch.cat((image_x, txt_x), dim=1))
        return self.loss_fct(x.view(-1, 2), label.view(-1))
