"""fp32 parity mode of the inference path (BASELINE.json north_star: "In fp32 mode, logits match within 1e-4
relative and greedy decodes match token-for-token").

Selected with `config.kmb_precision = "fp32"` (or env KMBART_PRECISION=fp32); the default is the bf16 tensor-core
path.  Same reference call stack as the fast path (src/model/modules.py:104-165 encoder, HF-3.0.2 BartDecoder /
DecoderLayer / SelfAttention incl. the cached step, src/model/model.py:397 LM head), same C-ABI library, but:
  * every Linear is a 3xTF32 tcgen05 GEMM: fp32 operands are split on device into tf32 hi/lo parts
    (kmb_split_tf32: A -> [hi|hi|lo], W -> [hi|lo|hi]) so hi*hi + hi*lo + lo*hi accumulates to fp32 accuracy in TMEM;
  * activations stay fp32 end to end (GEMM epilogues add bias / exact-erf GELU / residual in fp32);
  * attention runs in the fp32 instantiation of the decode-attention kernel (kmb_attn_f32).
It is a correctness mode: launches are issued eagerly, nothing is fused for speed, and training is not offered."""
import math

import torch

from . import lib as L
from .engine import Plan, F32, _ptr


class Fp32Path:
    def __init__(self, eng):
        self.eng = eng
        self.lib = eng.lib
        self._w3 = {}
        self._w3_version = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return self.eng.stream()

    def _split(self, x, side):
        rows, K = x.shape
        out = torch.empty(rows, 3 * K, dtype=F32, device=x.device)
        L.check(self.lib.kmb_split_tf32(x.data_ptr(), x.stride(0), out.data_ptr(), rows, K, side, self._stream()), "kmb_split_tf32")
        return out

    def _weight3(self, key, w):
        st = self.eng.store
        v = st.version()
        if v != self._w3_version:
            self._w3, self._w3_version = {}, v
        t = self._w3.get(key)
        if t is None:
            t = self._split(w.contiguous(), 1)
            self._w3[key] = t
        return t

    def linear(self, x, wkey, w, bias=None, act=L.ACT_NONE, residual=None):
        """x [M, K] fp32 contiguous, w [N, K] fp32 -> fp32 [M, N] = act(x w^T + bias) (+ residual)"""
        M, K = x.shape
        N = w.shape[0]
        a3, w3 = self._split(x, 0), self._weight3(wkey, w)
        out = torch.empty(M, N, dtype=F32, device=x.device)
        plan = Plan()
        plan.stream = self._stream()
        self.eng.gemm(plan, a3, w3, M, N, 3 * K, 3 * K, 3 * K, elt=1, bias=bias, act=act, residual=residual, ld_res=N, out_f32=out, ld_f32=N)
        plan.run()
        return out

    def layer_norm(self, x, gname):
        st = self.eng.store
        M, d = x.shape
        out = torch.empty_like(x)
        L.check(self.lib.kmb_layernorm_fwd(0, x.data_ptr(), _ptr(st.p32(gname + ".weight")), _ptr(st.p32(gname + ".bias")), 0, out.data_ptr(),
                                           0, 0, 0, M, d, 0.0, 0, 0, self._stream()), "kmb_layernorm_fwd")
        return out

    def attention(self, q, q_rs, k, v, kv_ss, kv_ps, row_div, pad_u8, pad_ld, rows, H, T, causal_mod, d):
        o = torch.empty(rows, d, dtype=F32, device=self.eng.device)
        L.check(self.lib.kmb_attn_f32(q, q_rs, k, v, kv_ss, kv_ps, row_div, _ptr(pad_u8), pad_ld, o.data_ptr(), d, rows, H, T, 64,
                                      causal_mod, 0.125, self._stream()), "kmb_attn_f32")
        return o

    def _p(self, name):
        return self.eng.store.p32(self.eng.n(name))

    def _self_block(self, lp, x, B, S, H, pad_u8, causal):
        d = x.shape[1]
        st = self.eng.store
        wq = st.fused32(lp + ".self_attn.q_proj.weight", 3).view(3 * d, d)
        qkv = self.linear(x, lp + ".qkv", wq, bias=st.fused32(lp + ".self_attn.q_proj.bias", 3))
        base = qkv.data_ptr()
        ctx = self.attention(base, 3 * d, base + 4 * d, base + 8 * d, S * 3 * d, 3 * d, S, pad_u8, S, B * S, H, S, S if causal else 0, d)
        pre = self.linear(ctx, lp + ".self_attn.out_proj", st.p32(lp + ".self_attn.out_proj.weight"),
                          bias=st.p32(lp + ".self_attn.out_proj.bias"), residual=x)
        return self.layer_norm(pre, lp + ".self_attn_layer_norm"), qkv

    def _ffn_block(self, lp, x):
        st = self.eng.store
        h = self.linear(x, lp + ".fc1", st.p32(lp + ".fc1.weight"), bias=st.p32(lp + ".fc1.bias"), act=L.ACT_GELU)
        pre = self.linear(h, lp + ".fc2", st.p32(lp + ".fc2.weight"), bias=st.p32(lp + ".fc2.bias"), residual=x)
        return self.layer_norm(pre, lp + ".final_layer_norm")

    def _cross_kv(self, lp, enc2d):
        st, d = self.eng.store, enc2d.shape[1]
        wkv = st.fused32(lp + ".encoder_attn.k_proj.weight", 2).view(2 * d, d)
        return self.linear(enc2d, lp + ".cross_kv", wkv, bias=st.fused32(lp + ".encoder_attn.k_proj.bias", 2))

    def _cross_block(self, lp, x, kv2, rows, row_div, Se, H, pad_e):
        st, d = self.eng.store, x.shape[1]
        q2 = self.linear(x, lp + ".encoder_attn.q_proj", st.p32(lp + ".encoder_attn.q_proj.weight"), bias=st.p32(lp + ".encoder_attn.q_proj.bias"))
        ctx = self.attention(q2.data_ptr(), d, kv2.data_ptr(), kv2.data_ptr() + 4 * d, Se * 2 * d, 2 * d, row_div, pad_e, Se, rows, H, Se, 0, d)
        pre = self.linear(ctx, lp + ".encoder_attn.out_proj", st.p32(lp + ".encoder_attn.out_proj.weight"),
                          bias=st.p32(lp + ".encoder_attn.out_proj.bias"), residual=x)
        return self.layer_norm(pre, lp + ".encoder_attn_layer_norm")

    def _embed(self, ids, S, which, slot=None, vis=None, boxes=None, w_box=None, position=None):
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        st = eng.store
        M = ids.numel()
        out = torch.empty(M, d, dtype=F32, device=eng.device)
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        pos_off = cfg.extra_pos_embeddings + (int(position) if position is not None else 0)
        L.check(self.lib.kmb_embed_ln_fwd(ids.data_ptr(), _ptr(slot), _ptr(st.p32(eng.n("shared.weight"))),
                                          _ptr(st.p32(eng.n(which + ".embed_positions.weight"))), _ptr(vis), _ptr(boxes), _ptr(w_box),
                                          _ptr(st.p32(eng.n("encoder.embed_images.linear.bias"))) if slot is not None else 0,
                                          _ptr(st.p32(eng.n(which + ".layernorm_embedding.weight"))),
                                          _ptr(st.p32(eng.n(which + ".layernorm_embedding.bias"))), 0, out.data_ptr(), 0, 0, 0, M,
                                          1 if position is not None else S, d, pos_off, 0, scale, 0.0, 0, 0, self._stream()), "kmb_embed_ln_fwd")
        return out

    # ------------------------------------------------------------------ public
    def encoder(self, input_ids, image_features, attention_mask):
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        dev = eng.device
        B, S = input_ids.shape
        ids = input_ids.reshape(-1).contiguous()
        counts = [int(f.shape[0]) for f in image_features]
        R = sum(counts)
        off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0).tolist()) if counts else [0], dtype=torch.int32).to(dev)
        slot = torch.empty(B * S, dtype=torch.int32, device=dev)
        L.check(self.lib.kmb_slot_index(ids.data_ptr(), off.data_ptr(), B, S, cfg.img_feat_id, cfg.cls_token_id, slot.data_ptr(),
                                        self._stream()), "kmb_slot_index")
        vis = boxes = w_box = None
        W = eng.store.p32(eng.n("encoder.embed_images.linear.weight"))
        if R > 0:
            allf = torch.cat([f for f in image_features if f.shape[0] > 0], 0).to(F32)
            feats, boxes = allf[:, :eng.fin - 4].contiguous(), allf[:, eng.fin - 4:].contiguous()
            vis = self.linear(feats, "img.feat", W[:, :eng.fin - 4])
            w_box = W[:, eng.fin - 4:].contiguous()
        else:
            vis = torch.zeros(1, d, dtype=F32, device=dev)
            boxes, w_box = torch.zeros(1, 4, dtype=F32, device=dev), W[:, eng.fin - 4:].contiguous()
        x = self._embed(ids, S, "encoder", slot=slot, vis=vis, boxes=boxes, w_box=w_box)
        pad = attention_mask.eq(0).to(torch.uint8).contiguous() if attention_mask is not None else None
        H = cfg.encoder_attention_heads
        for l in range(cfg.encoder_layers):
            lp = eng.n(f"encoder.layers.{l}")
            x, _ = self._self_block(lp, x, B, S, H, pad, False)
            x = self._ffn_block(lp, x)
        return x.view(B, S, d)

    def decoder_full(self, enc, attention_mask, decoder_input_ids, decoder_attention_mask):
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        B, Sd = decoder_input_ids.shape
        Se = enc.shape[1]
        ids = decoder_input_ids.reshape(-1).contiguous()
        if decoder_attention_mask is not None:
            pad_d = decoder_attention_mask.eq(0).to(torch.uint8).contiguous()
        else:
            pad_d = decoder_input_ids.eq(cfg.pad_token_id).to(torch.uint8).contiguous()
        pad_e = attention_mask.eq(0).to(torch.uint8).contiguous() if attention_mask is not None else None
        x = self._embed(ids, Sd, "decoder")
        enc2d = enc.reshape(B * Se, d).to(F32).contiguous()
        H = cfg.decoder_attention_heads
        for l in range(cfg.decoder_layers):
            lp = eng.n(f"decoder.layers.{l}")
            x, _ = self._self_block(lp, x, B, Sd, H, pad_d, True)
            x = self._cross_block(lp, x, self._cross_kv(lp, enc2d), B * Sd, Sd, Se, H, pad_e)
            x = self._ffn_block(lp, x)
        return x.view(B, Sd, d)

    def decoder_step(self, last_ids, position, enc, enc_pad_u8, caches):
        """One cached step; caches keep the legacy dict structure, tensors are fp32 [rows, heads, T, 64]."""
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        st = eng.store
        n = last_ids.shape[0]
        H = cfg.decoder_attention_heads
        ids = last_ids.reshape(-1).contiguous()
        x = self._embed(ids, 1, "decoder", position=position)
        new_caches = []
        enc2d = None
        for l in range(cfg.decoder_layers):
            lp = eng.n(f"decoder.layers.{l}")
            lc = caches[l] if caches is not None else {}
            wq = st.fused32(lp + ".self_attn.q_proj.weight", 3).view(3 * d, d)
            qkv = self.linear(x, lp + ".qkv", wq, bias=st.fused32(lp + ".self_attn.q_proj.bias", 3))
            k_new, v_new = qkv[:, d:2 * d].reshape(n, 1, d), qkv[:, 2 * d:].reshape(n, 1, d)
            sc = lc.get("self")
            if sc is not None and sc.get("prev_key") is not None:   # legacy layout [n, H, T, 64] -> token-major [n, T, d]
                pk = sc["prev_key"].permute(0, 2, 1, 3).reshape(n, -1, d)
                pv = sc["prev_value"].permute(0, 2, 1, 3).reshape(n, -1, d)
                K, V = torch.cat([pk, k_new], 1).contiguous(), torch.cat([pv, v_new], 1).contiguous()
            else:
                K, V = k_new.contiguous(), v_new.contiguous()
            T = K.shape[1]
            ctx = self.attention(qkv.data_ptr(), 3 * d, K.data_ptr(), V.data_ptr(), T * d, d, 1, None, 0, n, H, T, 0, d)
            pre = self.linear(ctx, lp + ".self_attn.out_proj", st.p32(lp + ".self_attn.out_proj.weight"),
                              bias=st.p32(lp + ".self_attn.out_proj.bias"), residual=x)
            y = self.layer_norm(pre, lp + ".self_attn_layer_norm")
            cc = lc.get("encoder_decoder")
            if cc is not None and cc.get("prev_key") is not None:
                K2 = cc["prev_key"].permute(0, 2, 1, 3).reshape(n, -1, d).contiguous()
                V2 = cc["prev_value"].permute(0, 2, 1, 3).reshape(n, -1, d).contiguous()
            else:
                Se = enc.shape[1]
                if enc2d is None:
                    enc2d = enc.reshape(n * Se, d).to(F32).contiguous()
                kv2 = self._cross_kv(lp, enc2d).view(n, Se, 2 * d)
                K2, V2 = kv2[:, :, :d].contiguous(), kv2[:, :, d:].contiguous()
            Se = K2.shape[1]
            q2 = self.linear(y, lp + ".encoder_attn.q_proj", st.p32(lp + ".encoder_attn.q_proj.weight"), bias=st.p32(lp + ".encoder_attn.q_proj.bias"))
            ctx2 = self.attention(q2.data_ptr(), d, K2.data_ptr(), V2.data_ptr(), Se * d, d, 1, enc_pad_u8, Se, n, H, Se, 0, d)
            pre2 = self.linear(ctx2, lp + ".encoder_attn.out_proj", st.p32(lp + ".encoder_attn.out_proj.weight"),
                               bias=st.p32(lp + ".encoder_attn.out_proj.bias"), residual=y)
            z = self.layer_norm(pre2, lp + ".encoder_attn_layer_norm")
            x = self._ffn_block(lp, z)
            as_legacy = lambda t: t.view(n, -1, H, 64).permute(0, 2, 1, 3)
            new_caches.append({"self": {"prev_key": as_legacy(K), "prev_value": as_legacy(V), "prev_key_padding_mask": None},
                               "encoder_decoder": {"prev_key": as_legacy(K2), "prev_value": as_legacy(V2), "prev_key_padding_mask": None}})
        return x, new_caches

    def logits(self, h2d, final_logits_bias):
        E = self.eng.store.p32(self.eng.n("shared.weight"))
        return self.linear(h2d.to(F32).contiguous(), "lm_head", E, bias=final_logits_bias.reshape(-1).to(F32).contiguous())
