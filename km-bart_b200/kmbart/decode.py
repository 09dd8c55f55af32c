"""KV-cached decode chain — the hot loop of GenerationMixin.generate (reference
src/model/mixins.py:336-382 -> HF-3.0.2 _generate_no_beam_search / _generate_beam_search, one
`self(**model_inputs)` per token with torch.cat-grown caches and a per-step index_select of
every cached tensor, :419-434).

A DecodeSession owns, for one (batch, source length, rows, max_length) shape:
  * the cross-attention K/V of every decoder layer, computed ONCE per sample ([B*Se, 2d] bf16) and
    shared by all beams / return sequences of that sample (the reference replicates the encoder
    states per beam and re-selects the cross K/V every step);
  * a preallocated self-attention cache [rows, max_len, 3d] per layer; the fused QKV GEMM of step t
    stores q|k|v of the new token straight into position t (no copies, no cat);
  * a beam ancestry table slot_tbl[row, pos]: re-ordering beams permutes 4-byte table entries
    instead of moving the cache;
  * one CUDA graph per step t holding the whole chain embed -> 6 x (QKV, attention, out-proj, LN,
    cross-q, cross-attention, out-proj, LN, fc1+GELU, fc2, LN) -> LM head, plus — for greedy and
    top-k sampling — the token selection and the finished-sentence bookkeeping, so a step is ONE
    graph launch with no host synchronisation.
Per-step algorithmic HBM bytes: decoder + LM-head bf16 weights (176 MB for the base model) + the
self cache read so far + the cross K/V; `step_bytes()` reports them for bench.py."""
import ctypes
import gc
import math
import os

import torch

from . import lib as L
from .engine import Plan, BF16, F32, _ptr


def packed_decoder_weights(eng):
    """Ring-slot images of the decoder weights (kmb_decode_pack_weights), rebuilt when the bf16 shadow changes.
    Returns (buffer, bytes per layer, section offsets)."""
    cfg, d, st = eng.cfg, eng.cfg.d_model, eng.store
    cached = getattr(eng, "_decode_packed", None)
    if cached is not None and cached[0] == st.shadow_version:
        return cached[1:]
    H, F_, Ld = cfg.decoder_attention_heads, cfg.decoder_ffn_dim, cfg.decoder_layers
    off = (ctypes.c_int64 * 7)()
    L.check(eng.lib.kmb_decode_pack_offsets(d, H, F_, off), "kmb_decode_pack_offsets")
    per_layer = int(off[6])
    buf = cached[1] if cached is not None else torch.empty(Ld * per_layer, dtype=torch.uint8, device=eng.device)
    for l in range(Ld):
        lp, y = eng.n(f"decoder.layers.{l}"), L.DecodeLayer()
        y.w_qkv = _ptr(st.p16(lp + ".self_attn.q_proj.weight", 3 * d))
        for key, name in (("o", "self_attn.out_proj"), ("cq", "encoder_attn.q_proj"), ("co", "encoder_attn.out_proj"),
                          ("fc1", "fc1"), ("fc2", "fc2")):
            setattr(y, "w_" + key, _ptr(st.p16(f"{lp}.{name}.weight")))
        L.check(eng.lib.kmb_decode_pack_weights(ctypes.byref(y), d, H, F_, buf.data_ptr() + l * per_layer, eng.stream()),
                "kmb_decode_pack_weights")
    eng._decode_packed = (st.shadow_version, buf, per_layer, [int(o) for o in off])
    return eng._decode_packed[1:]


class BeamController:
    """Device-resident beam-search state + the two-kernel step (kmb_beam_step, csrc/decode_control.cu): the body of
    HF-3.0.2 _generate_beam_search (reached from src/model/mixins.py:336-366) without a host round trip per token.
    Finished hypotheses stay on the device; finalize() reads them back once and runs the reference's selection /
    padding epilogue on the host."""

    def __init__(self, dev, batch, num_beams, V, max_len, eos, pad, early_stopping, length_penalty, slot_tbl=None, ids_next=None):
        rows, K = batch * num_beams, 2 * num_beams
        self.dev, self.batch, self.num_beams, self.V, self.max_len, self.rows = dev, batch, num_beams, V, max_len, rows
        self.eos, self.pad, self.early_stopping, self.length_penalty = eos, pad, bool(early_stopping), float(length_penalty)
        i32, f32, f64 = torch.int32, torch.float32, torch.float64
        self.cand_val = torch.empty(rows, K, dtype=f32, device=dev)
        self.cand_tok = torch.empty(rows, K, dtype=i32, device=dev)
        self.beam_scores = torch.zeros(rows, dtype=f32, device=dev)
        self.hist = torch.zeros(rows, max_len, dtype=i32, device=dev)
        self.ids_next = ids_next if ids_next is not None else torch.zeros(rows, dtype=torch.int64, device=dev)
        self.beam_idx = torch.zeros(rows, dtype=i32, device=dev)
        self.done = torch.zeros(batch, dtype=i32, device=dev)
        self.hyp_n = torch.zeros(batch, dtype=i32, device=dev)
        self.hyp_score = torch.zeros(batch, num_beams + 1, dtype=f64, device=dev)
        self.hyp_len = torch.zeros(batch, num_beams + 1, dtype=i32, device=dev)
        self.hyp_tok = torch.zeros(batch, num_beams + 1, max_len, dtype=i32, device=dev)
        self.worst = torch.zeros(batch, dtype=f64, device=dev)
        self.done_count = torch.zeros(1, dtype=i32, device=dev)
        self.init_scores = torch.zeros(batch, num_beams, dtype=f32, device=dev)
        self.init_scores[:, 1:] = -1e9          # HF: beam_scores[:, 1:] = -1e9 when do_sample is False
        st = L.BeamState()
        st.batch, st.num_beams, st.K, st.V, st.max_len = batch, num_beams, K, V, max_len
        st.eos = -1 if eos is None else int(eos)
        st.pad = 0 if pad is None else int(pad)
        st.early_stopping, st.length_penalty = int(self.early_stopping), self.length_penalty
        st.cand_val, st.cand_tok, st.beam_scores, st.hist = _ptr(self.cand_val), _ptr(self.cand_tok), _ptr(self.beam_scores), _ptr(self.hist)
        st.slot_tbl = _ptr(slot_tbl)
        st.ids_next, st.beam_idx, st.done, st.hyp_n = _ptr(self.ids_next), _ptr(self.beam_idx), _ptr(self.done), _ptr(self.hyp_n)
        st.hyp_score, st.hyp_len, st.hyp_tok, st.worst = _ptr(self.hyp_score), _ptr(self.hyp_len), _ptr(self.hyp_tok), _ptr(self.worst)
        st.done_count = _ptr(self.done_count)
        self.state = st

    def reset(self, decoder_start_token_id):
        self.beam_scores.copy_(self.init_scores.view(-1))
        self.hist.zero_()
        self.hist[:, 0] = decoder_start_token_id
        self.ids_next.fill_(decoder_start_token_id)
        self.done.zero_(); self.hyp_n.zero_(); self.done_count.zero_()
        self.worst.fill_(1e9)

    def step(self, lib, logits, cur_len, force_token, ban_eos, stream):
        L.check(lib.kmb_beam_step(logits.data_ptr(), logits.stride(0), ctypes.byref(self.state), int(cur_len), int(force_token), int(ban_eos),
                                  stream), "kmb_beam_step")

    def all_done(self):
        return int(self.done_count.item()) >= self.batch

    def finalize(self, hyp_cls, cur_len, num_return_sequences, max_length):
        """HF-3.0.2 _generate_beam_search epilogue: open beams of unfinished batch elements become hypotheses, the
        num_return_sequences best of every element are padded into one LongTensor."""
        nb = self.num_beams
        done = self.done.cpu().tolist()
        hyp_n = self.hyp_n.cpu().tolist()
        hyp_score, hyp_len = self.hyp_score.cpu().tolist(), self.hyp_len.cpu().tolist()
        hyp_tok = self.hyp_tok.cpu()
        worst = self.worst.cpu().tolist()
        hist = self.hist.cpu().long()
        scores = self.beam_scores.cpu().tolist()
        best, lengths = [], []
        for b in range(self.batch):
            h = hyp_cls(nb, max_length, self.length_penalty, early_stopping=self.early_stopping)
            h.beams = [(hyp_score[b][i], hyp_tok[b, i, :hyp_len[b][i]].long()) for i in range(hyp_n[b])]
            h.worst_score = worst[b]
            if not done[b]:
                for i in range(nb):
                    h.add(hist[b * nb + i, :cur_len], scores[b * nb + i])
            ranked = sorted(h.beams, key=lambda x: x[0])
            for _ in range(num_return_sequences):
                hyp = ranked.pop()[1]
                lengths.append(len(hyp))
                best.append(hyp)
        if min(lengths) != max(lengths):
            assert self.pad is not None, "`Pad_token_id` has to be defined"
            width = min(max(lengths) + 1, max_length)
            decoded = torch.full((len(best), width), self.pad, dtype=torch.long)
            for i, hyp in enumerate(best):
                decoded[i, :lengths[i]] = hyp
                if lengths[i] < max_length:
                    decoded[i, lengths[i]] = self.eos
        else:
            decoded = torch.stack(best).long()
        return decoded.to(self.dev)


class DecodeSession:
    def __init__(self, eng, B, Se, rows, max_len, has_pad):
        cfg, dev, d = eng.cfg, eng.device, eng.cfg.d_model
        self.eng, self.B, self.Se, self.rows, self.max_len, self.has_pad = eng, B, Se, rows, max_len, has_pad
        assert rows % B == 0
        self.row_div = rows // B
        Ld, F_, V = cfg.decoder_layers, cfg.decoder_ffn_dim, cfg.vocab_size
        self.enc_b16 = torch.empty(B * Se, d, dtype=BF16, device=dev)
        self.pad_u8 = torch.zeros(B, Se, dtype=torch.uint8, device=dev)
        self.kv2 = [torch.empty(B * Se, 2 * d, dtype=BF16, device=dev) for _ in range(Ld)]
        self.cache = [torch.zeros(rows, max_len, 3 * d, dtype=BF16, device=dev) for _ in range(Ld)]
        self.slot_tbl = torch.arange(rows, dtype=torch.int32, device=dev).view(rows, 1).repeat(1, max_len).contiguous()
        self.rows_i32 = torch.arange(rows, dtype=torch.int32, device=dev).view(rows, 1)
        self.ids = torch.zeros(rows, dtype=torch.int64, device=dev)           # token fed to the next step
        self.x_f32 = [torch.empty(rows, d, dtype=F32, device=dev) for _ in range(2)]
        self.x_b16 = [torch.empty(rows, d, dtype=BF16, device=dev) for _ in range(2)]
        self.y_f32, self.y_b16 = torch.empty(rows, d, dtype=F32, device=dev), torch.empty(rows, d, dtype=BF16, device=dev)
        self.z_f32, self.z_b16 = torch.empty(rows, d, dtype=F32, device=dev), torch.empty(rows, d, dtype=BF16, device=dev)
        self.ctx, self.q2, self.lin = (torch.empty(rows, d, dtype=BF16, device=dev) for _ in range(3))
        self.h = torch.empty(rows, F_, dtype=BF16, device=dev)
        self.logits = torch.empty(rows, V, dtype=F32, device=dev)
        self.graphs = {}          # (t, mode) -> torch.cuda.CUDAGraph
        # persistent single-launch step (csrc/decode_mega.cu); KMBART_DECODE_CHAIN=1 keeps the per-op launch chain
        self.mega = (d in (128, 768, 1024) and cfg.decoder_attention_heads * 64 == d and F_ % d == 0 and max_len <= 512 and Se <= 512
                     and Ld <= L.DECODE_MAX_LAYERS and os.environ.get("KMBART_DECODE_CHAIN", "0") != "1")
        # KMBART_DECODE_CLUSTER=1: the 4-CTA-cluster variant of the persistent step (csrc/decode_cluster.cu, 6 phases per layer,
        # K split inside a cluster, LayerNorm on load); correct and bit-reproducible but not faster yet (profiles/r02_decode_analysis.md)
        self.cluster = self.mega and os.environ.get("KMBART_DECODE_CLUSTER", "0") == "1"
        if self.mega:
            self.lin_f32 = torch.zeros(rows, d, dtype=F32, device=dev)
            self.grid_barrier = torch.zeros(2, dtype=torch.int64, device=dev)
        if self.cluster:   # fp32 pre-LayerNorm residual stream (ping-pong) + per-strip row statistics
            self.np_stats = d // (24 if d == 768 else 32 if d == 1024 else 8)
            self.y_pre = [torch.zeros(rows, d, dtype=F32, device=dev) for _ in range(2)]
            self.ln_stats = torch.zeros(1 + 3 * Ld, rows, self.np_stats, 2, dtype=F32, device=dev)
        self.use_tbl = False
        # greedy / sampling bookkeeping (device resident)
        self.out = torch.zeros(rows, max_len, dtype=torch.int64, device=dev)
        self.unfinished = torch.ones(rows, dtype=torch.int64, device=dev)
        self.sent_len = torch.zeros(rows, dtype=torch.int64, device=dev)
        self.sel = None
        self.launches_per_step = 0
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)    # multinomial stream of kmb_sample_select, redrawn per generate()
        self.beam_ctl = {}

    # ------------------------------------------------------------------ one-time per generate() call
    def begin(self, enc_hidden, attention_mask, decoder_start_token_id, use_tbl):
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        st = eng.store
        self.enc_b16.copy_(enc_hidden.reshape(self.B * self.Se, d))
        if self.has_pad:
            torch.eq(attention_mask.reshape(self.B, self.Se), 0, out=self.pad_u8.view(torch.bool))
        plan = Plan()
        plan.stream = eng.stream()
        for l in range(cfg.decoder_layers):
            lp = eng.n(f"decoder.layers.{l}")
            eng.gemm(plan, self.enc_b16, st.p16(lp + ".encoder_attn.k_proj.weight", 2 * d), self.B * self.Se, 2 * d, d, d, d,
                     bias=st.fused32(lp + ".encoder_attn.k_proj.bias", 2), out_bf16=self.kv2[l])
        plan.run()
        if self.cluster:
            packed_decoder_weights(eng)    # re-pack (in place: captured graphs stay valid) when the weights changed
        self.use_tbl = use_tbl
        if use_tbl:
            self.slot_tbl.copy_(self.rows_i32.expand(self.rows, self.max_len))
        self.seed.random_()     # torch's CUDA generator: torch.manual_seed() makes sampling reproducible
        self.ids.fill_(decoder_start_token_id)
        self.out.zero_()
        self.out[:, 0] = decoder_start_token_id
        self.unfinished.fill_(1)
        self.sent_len.fill_(self.max_len)

    # ------------------------------------------------------------------ the kernel chain of step t
    def _mega_args(self, t):
        """struct KmbDecodeStep / KmbDecodeStepC of step t (include/kmbart.h)."""
        eng, cfg, d = self.eng, self.eng.cfg, self.eng.cfg.d_model
        st = eng.store
        a = L.DecodeStepC() if self.cluster else L.DecodeStep()
        a.rows, a.d, a.H, a.F, a.L, a.t = self.rows, d, cfg.decoder_attention_heads, cfg.decoder_ffn_dim, cfg.decoder_layers, t
        a.max_len, a.Se, a.row_div, a.pos_row = self.max_len, self.Se, self.row_div, cfg.extra_pos_embeddings + t
        a.embed_scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        a.attn_scale = 0.125
        a.ids = _ptr(self.ids)
        a.tok_emb = _ptr(st.p32(eng.n("shared.weight")))
        a.pos_emb = _ptr(st.p32(eng.n("decoder.embed_positions.weight")))
        a.lne_g = _ptr(st.p32(eng.n("decoder.layernorm_embedding.weight")))
        a.lne_b = _ptr(st.p32(eng.n("decoder.layernorm_embedding.bias")))
        a.slot_tbl = _ptr(self.slot_tbl) if self.use_tbl else 0
        a.key_pad = _ptr(self.pad_u8) if self.has_pad else 0
        a.x_f32, a.x_b16, a.ctx, a.h = _ptr(self.x_f32[0]), _ptr(self.x_b16[0]), _ptr(self.ctx), _ptr(self.h)
        a.barrier = _ptr(self.grid_barrier)
        if self.cluster:
            a.flags = int(os.environ.get("KMBART_DECODE_FLAGS", "3"))   # L2 prefetches off: no gain measured (profiles/r02g)
            a.y0, a.y1, a.stats = _ptr(self.y_pre[0]), _ptr(self.y_pre[1]), _ptr(self.ln_stats)
            packed, per_layer, off = packed_decoder_weights(eng)
        else:
            a.lin, a.q2 = _ptr(self.lin_f32), _ptr(self.q2)
        for l in range(cfg.decoder_layers):
            lp, y = eng.n(f"decoder.layers.{l}"), a.layers[l]
            if self.cluster:
                for k in range(6):
                    y.packed[k] = packed.data_ptr() + l * per_layer + off[k]
            y.w_qkv = _ptr(st.p16(lp + ".self_attn.q_proj.weight", 3 * d))
            y.b_qkv = _ptr(st.fused32(lp + ".self_attn.q_proj.bias", 3))
            for key, name in (("o", "self_attn.out_proj"), ("cq", "encoder_attn.q_proj"), ("co", "encoder_attn.out_proj"),
                              ("fc1", "fc1"), ("fc2", "fc2")):
                setattr(y, "w_" + key, _ptr(st.p16(f"{lp}.{name}.weight")))
                setattr(y, "b_" + key, _ptr(st.p32(f"{lp}.{name}.bias")))
            for i, name in enumerate(("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"), 1):
                setattr(y, f"ln{i}_g", _ptr(st.p32(f"{lp}.{name}.weight")))
                setattr(y, f"ln{i}_b", _ptr(st.p32(f"{lp}.{name}.bias")))
            y.cache, y.cross_kv = _ptr(self.cache[l]), _ptr(self.kv2[l])
        return a

    def _emit_step(self, plan, t):
        eng, cfg, lib, d = self.eng, self.eng.cfg, self.eng.lib, self.eng.cfg.d_model
        st, rows, H, F_ = eng.store, self.rows, cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        V = cfg.vocab_size
        if self.mega:
            args = self._mega_args(t)
            plan.add(lib.kmb_decode_step_cluster if self.cluster else lib.kmb_decode_step, ctypes.byref(args), plan.stream, keep=args)
            # LM head at M <= 128 is a pure weight stream (77 MB): narrow single-CTA tiles cut the wave-quantisation
            # loss of 197 tiles of 256 columns on 148 SMs (2 waves, the second a third full)
            eng.gemm(plan, self.x_b16[0], st.p16(eng.n("shared.weight")), rows, V, d, d, d, bias=self.flb, out_f32=self.logits, ld_f32=V,
                     tile_n=int(os.environ.get("KMBART_LMHEAD_TILE", "128")) if rows <= 128 else 0)
            return
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        x_f32, x_b16 = self.x_f32[0], self.x_b16[0]
        plan.add(lib.kmb_embed_ln_fwd, _ptr(self.ids), 0, _ptr(st.p32(eng.n("shared.weight"))),
                 _ptr(st.p32(eng.n("decoder.embed_positions.weight"))), 0, 0, 0, 0,
                 _ptr(st.p32(eng.n("decoder.layernorm_embedding.weight"))), _ptr(st.p32(eng.n("decoder.layernorm_embedding.bias"))),
                 0, _ptr(x_f32), _ptr(x_b16), 0, 0, rows, 1, d, cfg.extra_pos_embeddings + t, 0, scale, 0.0, 0, 0, plan.stream)
        ml3 = self.max_len * 3 * d
        tbl = _ptr(self.slot_tbl) if self.use_tbl else 0
        pad = _ptr(self.pad_u8) if self.has_pad else 0
        for l in range(cfg.decoder_layers):
            lp = eng.n(f"decoder.layers.{l}")
            cache = self.cache[l]
            slot_t = cache[:, t, :]                       # [rows, 3d] view, row stride max_len*3d
            eng.gemm(plan, x_b16, st.p16(lp + ".self_attn.q_proj.weight", 3 * d), rows, 3 * d, d, d, d,
                     bias=st.fused32(lp + ".self_attn.q_proj.bias", 3), out_bf16=slot_t, ld_bf16=ml3)
            base = cache.data_ptr()
            plan.add(lib.kmb_decode_attn, slot_t.data_ptr(), ml3, base + 2 * d, base + 4 * d, ml3, 3 * d, tbl, self.max_len, 1, 0, 0,
                     _ptr(self.ctx), d, rows, H, t + 1, 64, 0.125, plan.stream)
            eng.gemm(plan, self.ctx, st.p16(lp + ".self_attn.out_proj.weight"), rows, d, d, d, d,
                     bias=st.p32(lp + ".self_attn.out_proj.bias"), out_bf16=self.lin)
            eng.ln_fwd(plan, self.lin, x_f32, lp + ".self_attn_layer_norm", None, self.y_f32, self.y_b16, None, None, rows)
            eng.gemm(plan, self.y_b16, st.p16(lp + ".encoder_attn.q_proj.weight"), rows, d, d, d, d,
                     bias=st.p32(lp + ".encoder_attn.q_proj.bias"), out_bf16=self.q2)
            kv = self.kv2[l]
            plan.add(lib.kmb_decode_attn, _ptr(self.q2), d, kv.data_ptr(), kv.data_ptr() + 2 * d, self.Se * 2 * d, 2 * d, 0, 0,
                     self.row_div, pad, self.Se, _ptr(self.ctx), d, rows, H, self.Se, 64, 0.125, plan.stream)
            eng.gemm(plan, self.ctx, st.p16(lp + ".encoder_attn.out_proj.weight"), rows, d, d, d, d,
                     bias=st.p32(lp + ".encoder_attn.out_proj.bias"), out_bf16=self.lin)
            eng.ln_fwd(plan, self.lin, self.y_f32, lp + ".encoder_attn_layer_norm", None, self.z_f32, self.z_b16, None, None, rows)
            eng.gemm(plan, self.z_b16, st.p16(lp + ".fc1.weight"), rows, F_, d, d, d, bias=st.p32(lp + ".fc1.bias"),
                     act=L.ACT_GELU, out_bf16=self.h)
            eng.gemm(plan, self.h, st.p16(lp + ".fc2.weight"), rows, d, F_, F_, F_, bias=st.p32(lp + ".fc2.bias"), out_bf16=self.lin)
            nxt = (l + 1) % 2
            eng.ln_fwd(plan, self.lin, self.z_f32, lp + ".final_layer_norm", None, self.x_f32[nxt], self.x_b16[nxt], None, None, rows)
            x_f32, x_b16 = self.x_f32[nxt], self.x_b16[nxt]
        eng.gemm(plan, x_b16, st.p16(eng.n("shared.weight")), rows, V, d, d, d, bias=self.flb, out_f32=self.logits, ld_f32=V)

    # ------------------------------------------------------------------ selection (device side, captured with the step)
    def _select_greedy_or_sample(self, t, sel):
        """HF-3.0.2 _generate_no_beam_search body for cur_len = t + 1 (src/model/mixins.py:368-382 dispatch)."""
        cur_len = t + 1
        logits = self.logits
        if not sel["do_sample"]:   # one fused kernel: EOS ban, argmax, pad for finished rows, append, bookkeeping
            eos = -1 if sel["eos"] is None else int(sel["eos"])
            pad_id = 0 if sel["pad"] is None else int(sel["pad"])
            L.check(self.eng.lib.kmb_greedy_select(logits.data_ptr(), logits.shape[1], self.rows, logits.shape[1], eos, pad_id,
                                                   int(eos >= 0 and cur_len < sel["min_length"]), cur_len, self.unfinished.data_ptr(),
                                                   self.sent_len.data_ptr(), self.out.data_ptr(), self.max_len, self.ids.data_ptr(),
                                                   torch.cuda.current_stream(self.eng.device).cuda_stream), "kmb_greedy_select")
            return
        # top-k (or unfiltered, top_k == 0) multinomial sampling: one kernel per step (csrc/decode_control.cu)
        eos = -1 if sel["eos"] is None else int(sel["eos"])
        pad_id = 0 if sel["pad"] is None else int(sel["pad"])
        L.check(self.eng.lib.kmb_sample_select(logits.data_ptr(), logits.shape[1], self.rows, logits.shape[1], float(sel["temperature"]),
                                               int(sel["top_k"]), float(sel.get("top_p", 1.0)), eos, pad_id,
                                               int(eos >= 0 and cur_len < sel["min_length"]), cur_len,
                                               self.seed.data_ptr(), self.unfinished.data_ptr(), self.sent_len.data_ptr(), self.out.data_ptr(),
                                               self.max_len, self.ids.data_ptr(), torch.cuda.current_stream(self.eng.device).cuda_stream),
                "kmb_sample_select")

    def beam_controller(self, num_beams, early_stopping, length_penalty, eos, pad):
        key = (num_beams, bool(early_stopping), float(length_penalty), eos, pad)
        ctl = self.beam_ctl.get(key)
        if ctl is None:
            ctl = BeamController(self.eng.device, self.rows // num_beams, num_beams, self.eng.cfg.vocab_size, self.max_len, eos, pad,
                                 early_stopping, length_penalty, slot_tbl=self.slot_tbl, ids_next=self.ids)
            self.beam_ctl[key] = ctl
        return ctl

    def step(self, t, flb, sel=None, beam=None):
        """Replay (or first capture) the graph of step t.  sel: greedy / sampling selection is part of the graph;
        beam = (BeamController, force_token, ban_eos): the device-side beam step is part of the graph; neither: model chain
        + LM head only (callers read self.logits)."""
        key = (t, None if sel is None else tuple(sorted(sel.items())), self.use_tbl,
               None if beam is None else (id(beam[0]), beam[1], beam[2]))

        def select():
            if sel is not None:
                self._select_greedy_or_sample(t, sel)
            if beam is not None:
                beam[0].step(self.eng.lib, self.logits, t + 1, beam[1], beam[2], torch.cuda.current_stream(self.eng.device).cuda_stream)
        g = self.graphs.get(key)
        if g is None:
            self.flb = flb.reshape(-1)
            plan = Plan()
            plan.stream = None
            # warm-up run on a side stream (sets function attributes, lets torch ops pick their workspaces)
            s = torch.cuda.Stream(device=self.eng.device)
            s.wait_stream(torch.cuda.current_stream(self.eng.device))
            snapshot = [self.ids.clone(), self.out.clone(), self.unfinished.clone(), self.sent_len.clone(), self.slot_tbl.clone()]
            saved = [self.ids, self.out, self.unfinished, self.sent_len, self.slot_tbl]
            if beam is not None:
                c = beam[0]
                saved += [c.beam_scores, c.hist, c.done, c.hyp_n, c.hyp_score, c.hyp_len, c.hyp_tok, c.worst, c.done_count]
                snapshot += [x.clone() for x in saved[5:]]
            with torch.cuda.stream(s):
                plan.stream = s.cuda_stream
                self._emit_step(plan, t)
                plan.run()
                select()
            torch.cuda.current_stream(self.eng.device).wait_stream(s)
            for dst, src in zip(saved, snapshot):
                dst.copy_(src)
            g = torch.cuda.CUDAGraph()
            # destructors of unrelated objects (e.g. the graphs of a discarded model) must not run inside the capture
            gc.collect()
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, capture_error_mode="relaxed"):
                    cap = Plan()
                    cap.stream = torch.cuda.current_stream(self.eng.device).cuda_stream
                    self._emit_step(cap, t)
                    cap.run()
                    select()
            finally:
                if gc_was_enabled:
                    gc.enable()
            self.graphs[key] = g
            self.launches_per_step = len(cap)
            self._keep = getattr(self, "_keep", []) + [cap]
        g.replay()

    def reorder(self, beam_idx, t):
        """Beam step bookkeeping: new row j continues parent beam_idx[j] — permute the ancestry table rows (positions
        <= t) instead of index_select-ing the cache (src/model/mixins.py:419-434)."""
        self.slot_tbl.copy_(self.slot_tbl.index_select(0, beam_idx))
        if t + 1 < self.max_len:
            self.slot_tbl[:, t + 1:] = self.rows_i32

    def step_bytes(self, t):
        """Algorithmic HBM bytes of step t: weights once + self cache so far + cross K/V + logits."""
        cfg, d = self.eng.cfg, self.eng.cfg.d_model
        F_, V, Ld = cfg.decoder_ffn_dim, cfg.vocab_size, cfg.decoder_layers
        weights = 2 * (Ld * (6 * d * d + 2 * d * F_) + V * d)
        self_kv = self.rows * Ld * 2 * (t + 1) * d * 2
        cross_kv = self.B * Ld * 2 * self.Se * d * 2
        return weights + self_kv + cross_kv


def get_session(eng, B, Se, rows, max_len, has_pad):
    key = ("dec", B, Se, rows, max_len, has_pad)
    s = eng.arenas.get(key)
    if s is None:
        s = DecodeSession(eng, B, Se, rows, max_len, has_pad)
        eng.remember(key, s)      # least-recently-used bound shared with the training workspaces (a session holds its graphs)
    else:
        eng.touch(key)
    return s
