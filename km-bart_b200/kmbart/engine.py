"""Execution engine behind the `src.model` drop-in classes.

It owns (a) the flat parameter / gradient / bf16-shadow buffers that the nn.Parameters of the
model are views of, (b) a workspace arena per input shape, and (c) *launch plans*: the ordered
list of C-ABI kernel calls that make up a forward or backward pass for that shape.  A plan is
built once (pointers are stable because the arena is persistent) and then replayed with
almost no Python work per launch, which also makes the step CUDA-graph capturable.

Arithmetic contract (what the plans compute) follows the reference call stack in
SURVEY.md §3.1: src/model/model.py:325-405 -> :39-103 -> src/model/modules.py:104-165 and
HF-3.0.2 BartDecoder / DecoderLayer / SelfAttention, post-LN, exact-erf GELU.
Layout choices are this implementation's own: token-major [B*S, d] activations, an fp32
residual stream next to bf16 GEMM operands, fused QKV / cross-KV projections.
"""
import ctypes as C
import math
import os
from collections import OrderedDict

import torch

from . import lib as L
from .feed import unwrap_features

BF16, F32 = torch.bfloat16, torch.float32


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _zero(t):
    t.zero_()
    return 0


class Plan:
    """Ordered kernel launches with pre-marshalled arguments."""

    def __init__(self):
        self.calls = []
        self.keep = []

    def add(self, fn, *args, keep=None):
        self.calls.append((fn, args))
        if keep is not None:
            self.keep.append(keep)

    def run(self):
        for fn, args in self.calls:
            rc = fn(*args)
            if rc != 0:
                L.check(rc, getattr(fn, "__name__", "kernel"))

    def __len__(self):
        return len(self.calls)

    # kernels launched per C-ABI call when it is more than one (bench.py's gpu_launches claim)
    _MULTI = {"kmb_ce_combine": 3, "kmb_embed_bwd": 2, "kmb_adamw_multi": 2}   # kmb_attn_bwd: 1 fused kernel for S <= 128

    def kernel_count(self):
        n = 0
        for fn, _ in self.calls:
            name = getattr(fn, "__name__", "")
            if name.startswith("kmb_"):
                n += self._MULTI.get(name, 1)
            elif fn is _zero:
                n += 1
        return n


# --------------------------------------------------------------------------- parameters
_BIG_SUFFIXES = ("q_proj.weight", "k_proj.weight", "v_proj.weight", "out_proj.weight", "fc1.weight", "fc2.weight",
                 "shared.weight", "dense.weight")


class ParamStore:
    """Flat fp32 master / fp32 grad / bf16 shadow buffers; model parameters become views.

    Order: first every tensor whose gradient is accumulated with atomics or partial writes
    (biases, LayerNorm, positions, image projection) — one memset clears them — then the
    GEMM-written matrices.  q/k/v (and cross k/v) projections are adjacent so the fused
    [3d, d] / [2d, d] operands are plain views."""

    def __init__(self, model):
        named = list(model.named_parameters())
        self.device = named[0][1].device
        self.names = [n for n, _ in named]
        self.params = dict(named)
        groups = {}
        for n in self.names:
            parts = n.split(".")
            if len(parts) >= 3 and parts[-2] in ("q_proj", "k_proj", "v_proj"):
                attn = ".".join(parts[:-2])
                kind = parts[-3]
                members = ("q_proj", "k_proj", "v_proj") if kind == "self_attn" else ("k_proj", "v_proj")
                if parts[-2] in members:
                    groups[n] = [f"{attn}.{m}.{parts[-1]}" for m in members]
        small, big, placed = [], [], set()
        for n in self.names:
            if n in placed:
                continue
            grp = groups.get(n, [n])
            dst = big if n.endswith(_BIG_SUFFIXES) and "embed_images" not in n else small
            for m in grp:
                dst.append(m)
                placed.add(m)
        # the tied embedding / LM-head matrix goes last: its gradient is fully overwritten by the LM-head
        # wgrad GEMM, everything before it is cleared by one memset and then only accumulated into
        tail = [n for n in big if n.endswith("shared.weight")]
        big = [n for n in big if not n.endswith("shared.weight")] + tail
        self.offsets, off = {}, 0
        order = small + big
        for i, n in enumerate(order):
            numel = self.params[n].numel()
            in_group_tail = n in groups and groups[n][0] != n
            if not in_group_tail:
                off = (off + 63) // 64 * 64
            else:
                assert numel % 8 == 0
            self.offsets[n] = off
            if i == len(small) - 1:
                self.small_end = off + numel
            off += numel
        if not small:
            self.small_end = 0
        self.total = (off + 63) // 64 * 64
        self.big_names = big
        self.zero_end = self.offsets[tail[0]] if tail else self.total
        self.P = torch.zeros(self.total, dtype=F32, device=self.device)
        self.G = torch.zeros(self.total, dtype=F32, device=self.device)
        self.P16 = torch.zeros(self.total, dtype=BF16, device=self.device)
        self.adopt()
        self.shadow_version = None

    def adopt(self):
        """(Re-)point every parameter at its slice of the flat buffer, preserving values."""
        with torch.no_grad():
            for n, p in self.params.items():
                o = self.offsets[n]
                view = self.P[o:o + p.numel()].view(p.shape)
                if p.data_ptr() != view.data_ptr():
                    view.copy_(p.data.to(device=self.device, dtype=F32))
                    p.data = view
                p._kmb_store = self

    def is_adopted(self):
        base = self.P.data_ptr()
        return all(p.data_ptr() == base + 4 * self.offsets[n] for n, p in self.params.items())

    def version(self):
        return sum(p._version for p in self.params.values())

    def p32(self, name):
        p = self.params[name]
        o = self.offsets[name]
        return self.P[o:o + p.numel()].view(p.shape)

    def p16(self, name, rows=None):
        p = self.params[name]
        o = self.offsets[name]
        if rows is None:
            return self.P16[o:o + p.numel()].view(p.shape)
        cols = p.shape[-1] if p.dim() == 2 else 1
        return self.P16[o:o + rows * cols].view(rows, cols) if p.dim() == 2 else self.P16[o:o + rows]

    def fused32(self, name, count):
        """fp32 view spanning `count` adjacent same-shape tensors starting at `name` (biases)."""
        p = self.params[name]
        o = self.offsets[name]
        return self.P[o:o + count * p.numel()]

    def g(self, name, count=1):
        p = self.params[name]
        o = self.offsets[name]
        if p.dim() == 2:
            return self.G[o:o + count * p.numel()].view(count * p.shape[0], p.shape[1])
        return self.G[o:o + count * p.numel()]

    def grad_view(self, name):
        p = self.params[name]
        o = self.offsets[name]
        return self.G[o:o + p.numel()].view(p.shape)


# --------------------------------------------------------------------------- engine
class Engine:
    def __init__(self, model, config, prefix="model."):
        L.require_b200()
        self.lib = L.load()
        self.cfg = config
        self.prefix = prefix          # "model." for the LM classes, "" for the bare MultiModalBartModel
        self.model = model
        self.store = ParamStore(model)
        self.device = self.store.device
        # workspaces ("arenas": activation stash + staging buffers of one input shape; decode sessions) and their launch
        # plans, least recently used first.  The reference's collator pads to the longest sequence of each batch
        # (src/data/collation.py:68-213), so shapes keep changing: at most KMBART_MAX_ARENAS workspaces stay resident, the
        # oldest is dropped (its memory returns to torch's caching allocator and is reused by the next one).
        self.plans = {}
        self.arenas = OrderedDict()
        self.max_arenas = max(2, int(os.environ.get("KMBART_MAX_ARENAS", "8")))
        self._fwd_serial = 0
        d = config.d_model
        assert d % 64 == 0 and d <= 1024, "d_model must be a multiple of 64 and <= 1024"
        assert d // config.encoder_attention_heads == 64 and d // config.decoder_attention_heads == 64, \
            "the attention kernels are specialised for head_dim 64 (bart-base / bart-large)"
        self.fin = config.image_feature_size
        assert self.fin == 2052, "image_feature_size must be 2048 + 4 (src/model/config.py:11)"
        dev = self.device
        self.w_feat16 = torch.zeros(d, self.fin - 4, dtype=BF16, device=dev)
        self.w_box = torch.zeros(d, 4, dtype=F32, device=dev)
        self.seed_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)
        self.upstream = torch.ones(1, dtype=F32, device=dev)
        self.last_train = None
        self.launches_last = 0
        self.grad_reducer = None   # kmbart.parallel.FlatGradReducer: all-reduce points inside the backward plan

    # ------------------------------------------------------------------ helpers
    def n(self, name):
        return self.prefix + name

    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def sync_shadow(self, force=False):
        """Refresh bf16 shadow weights when the fp32 masters changed outside the fused AdamW."""
        st = self.store
        if not st.is_adopted():
            st.adopt()
            force = True
        v = st.version()
        if force or v != st.shadow_version:
            L.check(self.lib.kmb_cast_bf16(st.P.data_ptr(), st.P16.data_ptr(), st.total, self.stream()), "cast")
            st.shadow_version = v
        w = st.p32(self.n("encoder.embed_images.linear.weight"))
        L.check(self.lib.kmb_repack_img_weight(w.data_ptr(), self.w_feat16.data_ptr(), self.w_box.data_ptr(),
                                               self.cfg.d_model, self.fin, self.stream()), "repack")

    def arena(self, key):
        a = self.arenas.get(key)
        if a is None:
            a = {}
            self.remember(key, a)
        return a

    def remember(self, key, a):
        """Register a workspace as most recently used; evict beyond the bound (workspace + its plans / graphs)."""
        self.arenas[key] = a
        self.arenas.move_to_end(key)
        while len(self.arenas) > self.max_arenas:
            old, _ = self.arenas.popitem(last=False)
            self.plans.pop(old, None)

    def touch(self, key):
        if key in self.arenas:
            self.arenas.move_to_end(key)

    def _seed_ptr(self):
        return getattr(self, "_cur_seed", self.seed).data_ptr()

    @staticmethod
    def buf(a, name, shape, dtype):
        t = a.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=a["__dev"])
            a[name] = t
        return t

    # ------------------------------------------------------------------ plan emitters
    def gemm(self, plan, A, B, M, N, K, lda, ldb, a_mn=0, b_mn=0, elt=0, tile_n=0, **kw):
        e = L.GemmEpilogue()
        e.mode = kw.get("mode", L.EPI_LINEAR)
        e.act = kw.get("act", L.ACT_NONE)
        e.alpha = kw.get("alpha", 1.0)
        e.accumulate = int(kw.get("accumulate", 0))
        e.bias = _ptr(kw.get("bias"))
        res = kw.get("residual")
        e.residual, e.ld_res = _ptr(res), kw.get("ld_res", N)
        aux = kw.get("aux")
        e.aux, e.ld_aux = _ptr(aux), kw.get("ld_aux", N)
        e.out_f32, e.ld_f32 = _ptr(kw.get("out_f32")), kw.get("ld_f32", N)
        e.out_bf16, e.ld_bf16 = _ptr(kw.get("out_bf16")), kw.get("ld_bf16", N)
        e.out_preact = _ptr(kw.get("out_preact"))
        e.dropout_p = kw.get("dropout_p", 0.0)
        e.dropout_tag = kw.get("dropout_tag", 0)
        e.dropout_seed = self._seed_ptr() if e.dropout_p > 0 else 0
        e.labels = _ptr(kw.get("labels"))
        e.ce_max, e.ce_sum = _ptr(kw.get("ce_max")), _ptr(kw.get("ce_sum"))
        e.ce_label_logit, e.ce_lse, e.ce_gscale = _ptr(kw.get("ce_label_logit")), _ptr(kw.get("ce_lse")), _ptr(kw.get("ce_gscale"))
        if os.environ.get("KMBART_DUMP_GEMMS"):   # profiling aid: one line per planned GEMM, in launch order
            print(f"KMB_GEMM {M} {N} {K} a_mn={a_mn} b_mn={b_mn} mode={e.mode} act={e.act} bias={int(bool(e.bias))} "
                  f"f32={int(bool(e.out_f32))} b16={int(bool(e.out_bf16))} acc={e.accumulate} res={int(bool(e.residual))} "
                  f"drop={e.dropout_p}", flush=True)
        plan.add(self.lib.kmb_gemm, _ptr(A), _ptr(B), M, N, K, lda, ldb, a_mn, b_mn, elt, C.byref(e), tile_n,
                 plan.stream, keep=e)

    def ln_fwd(self, plan, z_b16, residual, gname, pre, out_f32, out_b16, mean, rstd, M, drop=(0.0, 0)):
        """y = LN(residual + dropout(z)); z None -> plain LN(residual)."""
        st, d = self.store, self.cfg.d_model
        plan.add(self.lib.kmb_layernorm_fwd, _ptr(z_b16), _ptr(residual), _ptr(st.p32(gname + ".weight")),
                 _ptr(st.p32(gname + ".bias")), _ptr(pre), _ptr(out_f32), _ptr(out_b16), _ptr(mean), _ptr(rstd), M, d,
                 drop[0], drop[1], self._seed_ptr(), plan.stream)

    def ln_bwd(self, plan, dyA, dyB, pre, mean, rstd, gname, dpre, dz, dbias, M, drop_in=(0.0, 0), drop_out=(0.0, 0)):
        st, d = self.store, self.cfg.d_model
        plan.add(self.lib.kmb_layernorm_bwd, _ptr(dyA), _ptr(dyB), _ptr(pre), _ptr(mean), _ptr(rstd),
                 _ptr(st.p32(gname + ".weight")), _ptr(dpre), _ptr(dz), _ptr(st.g(gname + ".weight")),
                 _ptr(st.g(gname + ".bias")), _ptr(dbias), M, d, drop_in[0], drop_in[1], drop_out[0], drop_out[1],
                 self._seed_ptr(), plan.stream)

    def attn_fwd(self, plan, q, k, v, ldq, ldk, ldv, o, lse, pad, B, H, Sq, Sk, causal):
        plan.add(self.lib.kmb_attn_fwd, _ptr(q), _ptr(k), _ptr(v), ldq, ldk, ldv, _ptr(o), self.cfg.d_model, _ptr(lse),
                 _ptr(pad), B, H, Sq, Sk, 64, int(causal), 0.125, plan.stream)

    def attn_bwd(self, plan, q, k, v, ldq, ldk, ldv, o, do, lse, dscr, pad, dq, dk, dv, lddq, lddk, lddv, B, H, Sq, Sk, causal):
        d = self.cfg.d_model
        plan.add(self.lib.kmb_attn_bwd, _ptr(q), _ptr(k), _ptr(v), ldq, ldk, ldv, _ptr(o), d, _ptr(do), d, _ptr(lse),
                 _ptr(dscr), _ptr(pad), _ptr(dq), _ptr(dk), _ptr(dv), lddq, lddk, lddv, B, H, Sq, Sk, 64, int(causal),
                 0.125, plan.stream)

    def colsum(self, plan, x, ld, out, M, N):
        plan.add(self.lib.kmb_colsum_bf16, _ptr(x), ld, _ptr(out), M, N, plan.stream)

    # ------------------------------------------------------------------ transformer blocks (forward)
    def _self_block_fwd(self, plan, a, tag, lp, x_f32, x_b16, M, B, S, H, pad, causal, p_drop, drop_tag, train):
        """x -> LN(x + drop(out_proj(attn(qkv(x)))))   (post-LN, HF-3.0.2 Encoder/DecoderLayer first block)"""
        st, d = self.store, self.cfg.d_model
        qkv = self.buf(a, tag + "qkv", (M, 3 * d), BF16)
        ctx = self.buf(a, tag + "ctx", (M, d), BF16)
        lse = self.buf(a, tag + "lse", (B * H * S,), F32) if train else None
        pre = self.buf(a, tag + "pre1", (M, d), F32)
        y_f32 = self.buf(a, tag + "x1_f32", (M, d), F32)
        y_b16 = self.buf(a, tag + "x1_b16", (M, d), BF16)
        mean = self.buf(a, tag + "mean1", (M,), F32)
        rstd = self.buf(a, tag + "rstd1", (M,), F32)
        self.gemm(plan, x_b16, st.p16(lp + ".self_attn.q_proj.weight", 3 * d), M, 3 * d, d, d, d,
                  bias=st.fused32(lp + ".self_attn.q_proj.bias", 3), out_bf16=qkv)
        self.attn_fwd(plan, qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d, 3 * d, 3 * d, ctx, lse, pad, B, H, S, S, causal)
        z = self.buf(a, f"zbuf{M}", (M, d), BF16)
        self.gemm(plan, ctx, st.p16(lp + ".self_attn.out_proj.weight"), M, d, d, d, d,
                  bias=st.p32(lp + ".self_attn.out_proj.bias"), out_bf16=z)
        self.ln_fwd(plan, z, x_f32, lp + ".self_attn_layer_norm", pre, y_f32, y_b16, mean, rstd, M, drop=(p_drop, drop_tag))
        return y_f32, y_b16

    def _ffn_block_fwd(self, plan, a, tag, lp, x_f32, x_b16, M, F, p_drop, drop_tag, train):
        st, d = self.store, self.cfg.d_model
        u = self.buf(a, tag + "u", (M, F), BF16) if train else None
        h = self.buf(a, tag + "h", (M, F), BF16)
        pre = self.buf(a, tag + "pre3", (M, d), F32)
        y_f32 = self.buf(a, tag + "x3_f32", (M, d), F32)
        y_b16 = self.buf(a, tag + "x3_b16", (M, d), BF16)
        mean = self.buf(a, tag + "mean3", (M,), F32)
        rstd = self.buf(a, tag + "rstd3", (M,), F32)
        self.gemm(plan, x_b16, st.p16(lp + ".fc1.weight"), M, F, d, d, d, bias=st.p32(lp + ".fc1.bias"),
                  act=L.ACT_GELU, out_bf16=h, out_preact=u)
        z = self.buf(a, f"zbuf{M}", (M, d), BF16)
        self.gemm(plan, h, st.p16(lp + ".fc2.weight"), M, d, F, F, F, bias=st.p32(lp + ".fc2.bias"), out_bf16=z)
        self.ln_fwd(plan, z, x_f32, lp + ".final_layer_norm", pre, y_f32, y_b16, mean, rstd, M, drop=(p_drop, drop_tag))
        return y_f32, y_b16

    def _cross_block_fwd(self, plan, a, tag, lp, x_f32, x_b16, enc_b16, Md, Me, B, Sd, Se, H, pad_e, p_drop, drop_tag, train):
        st, d = self.store, self.cfg.d_model
        q2 = self.buf(a, tag + "q2", (Md, d), BF16)
        kv2 = self.buf(a, tag + "kv2", (Me, 2 * d), BF16)
        ctx = self.buf(a, tag + "ctx2", (Md, d), BF16)
        lse = self.buf(a, tag + "lse2", (B * H * Sd,), F32) if train else None
        pre = self.buf(a, tag + "pre2", (Md, d), F32)
        y_f32 = self.buf(a, tag + "x2_f32", (Md, d), F32)
        y_b16 = self.buf(a, tag + "x2_b16", (Md, d), BF16)
        mean = self.buf(a, tag + "mean2", (Md,), F32)
        rstd = self.buf(a, tag + "rstd2", (Md,), F32)
        self.gemm(plan, x_b16, st.p16(lp + ".encoder_attn.q_proj.weight"), Md, d, d, d, d,
                  bias=st.p32(lp + ".encoder_attn.q_proj.bias"), out_bf16=q2)
        self.gemm(plan, enc_b16, st.p16(lp + ".encoder_attn.k_proj.weight", 2 * d), Me, 2 * d, d, d, d,
                  bias=st.fused32(lp + ".encoder_attn.k_proj.bias", 2), out_bf16=kv2)
        self.attn_fwd(plan, q2, kv2, kv2[:, d:], d, 2 * d, 2 * d, ctx, lse, pad_e, B, H, Sd, Se, False)
        z = self.buf(a, f"zbuf{Md}", (Md, d), BF16)
        self.gemm(plan, ctx, st.p16(lp + ".encoder_attn.out_proj.weight"), Md, d, d, d, d,
                  bias=st.p32(lp + ".encoder_attn.out_proj.bias"), out_bf16=z)
        self.ln_fwd(plan, z, x_f32, lp + ".encoder_attn_layer_norm", pre, y_f32, y_b16, mean, rstd, Md, drop=(p_drop, drop_tag))
        return y_f32, y_b16

    # ------------------------------------------------------------------ forward plans
    def _encoder_fwd(self, plan, a, B, Se, R, train, p_drop):
        cfg, st, d = self.cfg, self.store, self.cfg.d_model
        Me = B * Se
        ids = a["ids_e"]
        pad_e = a["pad_e"] if a["has_mask_e"] else None
        slot = self.buf(a, "slot", (Me,), torch.int32)
        vis = self.buf(a, "vis_acc", (max(R, 1), d), F32)
        feats16 = self.buf(a, "feats16", (max(R, 1), self.fin - 4), BF16)
        boxes = self.buf(a, "boxes", (max(R, 1), 4), F32)
        if a["has_mask_e"]:
            plan.add(self.lib.kmb_invert_mask, _ptr(a["amask_e"]), _ptr(pad_e), Me, plan.stream)
        if R > 0:
            plan.add(self.lib.kmb_pack_features, _ptr(a["feat_ptrs"]) if a["packed"] is None else 0, _ptr(a["row_off"]), B,
                     _ptr(a["packed"]), _ptr(feats16), _ptr(boxes), R, plan.stream)
            self.gemm(plan, feats16, self.w_feat16, R, d, self.fin - 4, self.fin - 4, self.fin - 4, out_f32=vis)
        plan.add(self.lib.kmb_slot_index, _ptr(ids), _ptr(a["row_off"]), B, Se, cfg.img_feat_id, cfg.cls_token_id,
                 _ptr(slot), plan.stream)
        x_f32 = self.buf(a, "e.x0_f32", (Me, d), F32)
        x_b16 = self.buf(a, "e.x0_b16", (Me, d), BF16)
        emb_pre = self.buf(a, "e.emb_pre", (Me, d), F32) if train else None
        mean = self.buf(a, "e.emb_mean", (Me,), F32)
        rstd = self.buf(a, "e.emb_rstd", (Me,), F32)
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        plan.add(self.lib.kmb_embed_ln_fwd, _ptr(ids), _ptr(slot), _ptr(st.p32(self.n("shared.weight"))),
                 _ptr(st.p32(self.n("encoder.embed_positions.weight"))), _ptr(vis), _ptr(boxes), _ptr(self.w_box),
                 _ptr(st.p32(self.n("encoder.embed_images.linear.bias"))),
                 _ptr(st.p32(self.n("encoder.layernorm_embedding.weight"))),
                 _ptr(st.p32(self.n("encoder.layernorm_embedding.bias"))), _ptr(emb_pre), _ptr(x_f32), _ptr(x_b16),
                 _ptr(mean), _ptr(rstd), Me, Se, d, cfg.extra_pos_embeddings, 0, scale, p_drop, 1,
                 self._seed_ptr(), plan.stream)
        H, F = cfg.encoder_attention_heads, cfg.encoder_ffn_dim
        for l in range(cfg.encoder_layers):
            lp, tag = self.n(f"encoder.layers.{l}"), f"e{l}."
            a[tag + "in_f32"], a[tag + "in_b16"] = x_f32, x_b16
            x_f32, x_b16 = self._self_block_fwd(plan, a, tag, lp, x_f32, x_b16, Me, B, Se, H, pad_e, False, p_drop, 10 + 2 * l, train)
            x_f32, x_b16 = self._ffn_block_fwd(plan, a, tag, lp, x_f32, x_b16, Me, F, p_drop, 11 + 2 * l, train)
        a["enc_f32"], a["enc_b16"] = x_f32, x_b16
        return x_f32, x_b16

    def _decoder_fwd(self, plan, a, B, Sd, Se, train, p_drop, enc_b16):
        cfg, st, d = self.cfg, self.store, self.cfg.d_model
        Md, Me = B * Sd, B * Se
        ids = a["ids_d"]
        pad_d = a["pad_d"]
        pad_e = a["pad_e"] if a["has_mask_e"] else None
        x_f32 = self.buf(a, "d.x0_f32", (Md, d), F32)
        x_b16 = self.buf(a, "d.x0_b16", (Md, d), BF16)
        emb_pre = self.buf(a, "d.emb_pre", (Md, d), F32) if train else None
        mean = self.buf(a, "d.emb_mean", (Md,), F32)
        rstd = self.buf(a, "d.emb_rstd", (Md,), F32)
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        plan.add(self.lib.kmb_embed_ln_fwd, _ptr(ids), 0, _ptr(st.p32(self.n("shared.weight"))),
                 _ptr(st.p32(self.n("decoder.embed_positions.weight"))), 0, 0, 0, 0,
                 _ptr(st.p32(self.n("decoder.layernorm_embedding.weight"))),
                 _ptr(st.p32(self.n("decoder.layernorm_embedding.bias"))), _ptr(emb_pre), _ptr(x_f32), _ptr(x_b16),
                 _ptr(mean), _ptr(rstd), Md, Sd, d, cfg.extra_pos_embeddings, 0, scale, p_drop, 2,
                 self._seed_ptr(), plan.stream)
        H, F = cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        for l in range(cfg.decoder_layers):
            lp, tag = self.n(f"decoder.layers.{l}"), f"d{l}."
            a[tag + "in_f32"], a[tag + "in_b16"] = x_f32, x_b16
            x_f32, x_b16 = self._self_block_fwd(plan, a, tag, lp, x_f32, x_b16, Md, B, Sd, H, pad_d, True, p_drop, 100 + 3 * l, train)
            a[tag + "mid_f32"], a[tag + "mid_b16"] = x_f32, x_b16
            x_f32, x_b16 = self._cross_block_fwd(plan, a, tag, lp, x_f32, x_b16, enc_b16, Md, Me, B, Sd, Se, H, pad_e, p_drop, 101 + 3 * l, train)
            a[tag + "mid2_f32"], a[tag + "mid2_b16"] = x_f32, x_b16
            x_f32, x_b16 = self._ffn_block_fwd(plan, a, tag, lp, x_f32, x_b16, Md, F, p_drop, 102 + 3 * l, train)
        a["dec_f32"], a["dec_b16"] = x_f32, x_b16
        return x_f32, x_b16

    def _lm_loss_fwd(self, plan, a, Md, factor, add_total):
        cfg, st, d, V = self.cfg, self.store, self.cfg.d_model, self.cfg.vocab_size
        tile_n = 1256   # CTA-pair 256 x 256 tiles
        nt = self.lib.kmb_gemm_n_tiles(V, tile_n)
        ce_max = self.buf(a, "ce_max", (Md, nt), F32)
        ce_sum = self.buf(a, "ce_sum", (Md, nt), F32)
        lab_logit = self.buf(a, "ce_lab", (Md,), F32)
        lse = self.buf(a, "ce_lse", (Md,), F32)
        acc2 = self.buf(a, "ce_acc", (2,), F32)
        lm_loss = self.buf(a, "lm_loss", (1,), F32)
        plan.add(_zero, lab_logit)
        self.gemm(plan, a["dec_b16"], st.p16(self.n("shared.weight")), Md, V, d, d, d, tile_n=tile_n,
                  mode=L.EPI_CE_STATS, bias=a["flb"], labels=a["labels"], ce_max=ce_max, ce_sum=ce_sum,
                  ce_label_logit=lab_logit)
        plan.add(self.lib.kmb_ce_combine, _ptr(ce_max), _ptr(ce_sum), _ptr(lab_logit), _ptr(a["labels"]), Md, nt, _ptr(lse),
                 0, _ptr(acc2), float(factor), _ptr(lm_loss), _ptr(a["loss"]), int(add_total), plan.stream)

    # ------------------------------------------------------------------ backward plans
    # Backward convention: the gradient w.r.t. a block's output arrives as an fp32 part dyA (the
    # residual path: dpre of the following LayerNorm) plus a bf16 part dyB (dgrad GEMM output of the
    # consuming Linear); either may be None.  A block writes the fp32 part of its input gradient
    # into dpre_out (fresh buffer, distinct from dyA) and the bf16 part into dyb_out.
    def _ffn_block_bwd(self, plan, a, tag, lp, dyA, dyB, M, F, x_in_b16, p_drop, drop_tag, acc, dpre_out, dyb_out):
        st, d = self.store, self.cfg.d_model
        sfx = "" if M == a["Me"] else "_d"
        dz = self.buf(a, "g.dz" + sfx, (M, d), BF16)
        du = self.buf(a, "g.du" + sfx, (M, F), BF16)
        self.ln_bwd(plan, dyA, dyB, a[tag + "pre3"], a[tag + "mean3"], a[tag + "rstd3"], lp + ".final_layer_norm", dpre_out, dz,
                    st.g(lp + ".fc2.bias"), M, drop_out=(p_drop, drop_tag))
        # dW2[d, F] = dz^T h ; du = (dz W2) * gelu'(u) ; db1 = colsum(du) ; dW1[F, d] = du^T x ; dx = du W1
        self.gemm(plan, dz, a[tag + "h"], d, F, M, d, F, a_mn=1, b_mn=1, out_f32=st.g(lp + ".fc2.weight"), ld_f32=F, accumulate=acc)
        self.gemm(plan, dz, st.p16(lp + ".fc2.weight"), M, F, d, d, F, b_mn=1, act=L.ACT_GELU_GRAD, aux=a[tag + "u"], ld_aux=F, out_bf16=du)
        self.colsum(plan, du, F, st.g(lp + ".fc1.bias"), M, F)
        self.gemm(plan, du, x_in_b16, F, d, M, F, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".fc1.weight"), ld_f32=d, accumulate=acc)
        self.gemm(plan, du, st.p16(lp + ".fc1.weight"), M, d, F, F, d, b_mn=1, out_bf16=dyb_out)

    def _self_block_bwd(self, plan, a, tag, lp, dyA, dyB, M, B, S, H, pad, causal, x_in_b16, p_drop, drop_tag, acc, dpre_out, dyb_out):
        st, d = self.store, self.cfg.d_model
        sfx = "" if M == a["Me"] else "_d"
        dz = self.buf(a, "g.dz" + sfx, (M, d), BF16)
        dctx = self.buf(a, "g.dctx" + sfx, (M, d), BF16)
        dqkv = self.buf(a, "g.dqkv" + sfx, (M, 3 * d), BF16)
        dscr = self.buf(a, "g.dscr" + sfx, (B * H * S,), F32)
        qkv = a[tag + "qkv"]
        self.ln_bwd(plan, dyA, dyB, a[tag + "pre1"], a[tag + "mean1"], a[tag + "rstd1"], lp + ".self_attn_layer_norm", dpre_out, dz,
                    st.g(lp + ".self_attn.out_proj.bias"), M, drop_out=(p_drop, drop_tag))
        self.gemm(plan, dz, a[tag + "ctx"], d, d, M, d, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".self_attn.out_proj.weight"), ld_f32=d, accumulate=acc)
        self.gemm(plan, dz, st.p16(lp + ".self_attn.out_proj.weight"), M, d, d, d, d, b_mn=1, out_bf16=dctx)
        self.attn_bwd(plan, qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d, 3 * d, 3 * d, a[tag + "ctx"], dctx, a[tag + "lse"], dscr, pad,
                      dqkv, dqkv[:, d:], dqkv[:, 2 * d:], 3 * d, 3 * d, 3 * d, B, H, S, S, causal)
        self.colsum(plan, dqkv, 3 * d, st.g(lp + ".self_attn.q_proj.bias", 3), M, 3 * d)
        self.gemm(plan, dqkv, x_in_b16, 3 * d, d, M, 3 * d, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".self_attn.q_proj.weight", 3), ld_f32=d, accumulate=acc)
        self.gemm(plan, dqkv, st.p16(lp + ".self_attn.q_proj.weight", 3 * d), M, d, 3 * d, 3 * d, d, b_mn=1, out_bf16=dyb_out)

    def _cross_block_bwd(self, plan, a, tag, lp, dyA, dyB, Md, Me, B, Sd, Se, H, pad_e, x_in_b16, enc_b16, p_drop, drop_tag, acc,
                         dpre_out, dyb_out, denc, first_cross):
        st, d = self.store, self.cfg.d_model
        dz = self.buf(a, "g.dz_d", (Md, d), BF16)
        dctx = self.buf(a, "g.dctx_d", (Md, d), BF16)
        dq2 = self.buf(a, "g.dq2", (Md, d), BF16)
        dkv2 = self.buf(a, "g.dkv2", (Me, 2 * d), BF16)
        dscr = self.buf(a, "g.dscr_d", (B * H * Sd,), F32)
        q2, kv2 = a[tag + "q2"], a[tag + "kv2"]
        self.ln_bwd(plan, dyA, dyB, a[tag + "pre2"], a[tag + "mean2"], a[tag + "rstd2"], lp + ".encoder_attn_layer_norm", dpre_out, dz,
                    st.g(lp + ".encoder_attn.out_proj.bias"), Md, drop_out=(p_drop, drop_tag))
        self.gemm(plan, dz, a[tag + "ctx2"], d, d, Md, d, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".encoder_attn.out_proj.weight"), ld_f32=d, accumulate=acc)
        self.gemm(plan, dz, st.p16(lp + ".encoder_attn.out_proj.weight"), Md, d, d, d, d, b_mn=1, out_bf16=dctx)
        self.attn_bwd(plan, q2, kv2, kv2[:, d:], d, 2 * d, 2 * d, a[tag + "ctx2"], dctx, a[tag + "lse2"], dscr, pad_e,
                      dq2, dkv2, dkv2[:, d:], d, 2 * d, 2 * d, B, H, Sd, Se, False)
        self.colsum(plan, dq2, d, st.g(lp + ".encoder_attn.q_proj.bias"), Md, d)
        self.colsum(plan, dkv2, 2 * d, st.g(lp + ".encoder_attn.k_proj.bias", 2), Me, 2 * d)
        self.gemm(plan, dq2, x_in_b16, d, d, Md, d, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".encoder_attn.q_proj.weight"), ld_f32=d, accumulate=acc)
        self.gemm(plan, dkv2, enc_b16, 2 * d, d, Me, 2 * d, d, a_mn=1, b_mn=1, out_f32=st.g(lp + ".encoder_attn.k_proj.weight", 2), ld_f32=d, accumulate=acc)
        self.gemm(plan, dkv2, st.p16(lp + ".encoder_attn.k_proj.weight", 2 * d), Me, d, 2 * d, 2 * d, d, b_mn=1, out_f32=denc,
                  accumulate=0 if first_cross else 1)
        self.gemm(plan, dq2, st.p16(lp + ".encoder_attn.q_proj.weight"), Md, d, d, d, d, b_mn=1, out_bf16=dyb_out)

    def _build_train_fwd(self, a):
        cfg = self.cfg
        B, Se, Sd, R = a["B"], a["Se"], a["Sd"], a["R"]
        Me, Md = B * Se, B * Sd
        a["Me"], a["Md"] = Me, Md
        p_drop = float(cfg.dropout) if a["training"] else 0.0
        if a["training"] and (float(cfg.attention_dropout) > 0.0 or float(cfg.activation_dropout) > 0.0):
            raise ValueError("attention_dropout / activation_dropout > 0 are not implemented by the B200 kernels (only `dropout` is)")
        fwd = Plan()
        fwd.stream = a["stream"]
        # the dropout seed of a step lives in ITS workspace: a forward on another shape between this forward and its
        # backward (two micro-batches, loss_a + loss_b) must not change the masks the backward regenerates
        self._cur_seed = a["seed"]
        if p_drop > 0:
            fwd.add(self.lib.kmb_next_seed, _ptr(self.seed_state), _ptr(a["seed"]), fwd.stream)
        _, enc_b16 = self._encoder_fwd(fwd, a, B, Se, R, True, p_drop)
        self._decoder_fwd(fwd, a, B, Sd, Se, True, p_drop, enc_b16)
        if a["has_lm"]:
            self._lm_loss_fwd(fwd, a, Md, a["lm_factor"], add_total=False)
        else:
            fwd.add(_zero, a["loss"])
        if a["with_heads"]:   # fp32 gradient of the decoder output contributed by the pre-training heads
            fwd.add(_zero, self.buf(a, "g.dhead", (Md, self.cfg.d_model), F32))
        return fwd

    def _build_train_bwd(self, a, acc):
        """acc = 1: gradients are added onto the flat gradient buffer (gradient accumulation /
        zero_grad(set_to_none=False)); acc = 0: the buffer is overwritten."""
        cfg, st, d, V = self.cfg, self.store, self.cfg.d_model, self.cfg.vocab_size
        B, Se, Sd, R = a["B"], a["Se"], a["Sd"], a["R"]
        Me, Md = B * Se, B * Sd
        p_drop = float(cfg.dropout) if a["training"] else 0.0
        acc = int(acc)
        self._cur_seed = a["seed"]
        bwd = Plan()
        bwd.stream = a["stream"]
        has_lm, with_heads = a["has_lm"], a["with_heads"]
        assert has_lm or with_heads, "backward without any loss term"
        if not acc:   # (the heads' backward, which runs before this plan, clears the buffer itself: see heads.py)
            if not with_heads:
                bwd.add(_zero, st.G[:st.zero_end] if has_lm else st.G)
            elif not has_lm:
                bwd.add(_zero, st.G[st.zero_end:])
        lm_acc = acc
        acc = 1   # every other weight gradient accumulates onto the cleared buffer (enables split-K reds)
        dpre_d = [self.buf(a, "g.dpreA_d", (Md, d), F32), self.buf(a, "g.dpreB_d", (Md, d), F32)]
        dyb_d = self.buf(a, "g.dyb_d", (Md, d), BF16)
        if has_lm:
            gscale = self.buf(a, "ce_gscale", (1,), F32)
            dlog = self.buf(a, "dlogits", (Md, V), BF16)
            E16 = st.p16(self.n("shared.weight"))
            bwd.add(self.lib.kmb_ce_gscale, _ptr(a["ce_acc"]), _ptr(self.upstream), float(a["lm_factor"]), _ptr(gscale), bwd.stream)
            self.gemm(bwd, a["dec_b16"], E16, Md, V, d, d, d, tile_n=1256, mode=L.EPI_CE_GRAD, bias=a["flb"], labels=a["labels"],
                      ce_lse=a["ce_lse"], ce_gscale=gscale, out_bf16=dlog, ld_bf16=V)
            self.gemm(bwd, dlog, E16, Md, d, V, V, d, b_mn=1, out_bf16=dyb_d)
            self.gemm(bwd, dlog, a["dec_b16"], V, d, Md, V, d, a_mn=1, b_mn=1, out_f32=st.g(self.n("shared.weight")), ld_f32=d, accumulate=lm_acc)
        # decoder layers, last to first
        dyA, dyB, flip = (a["g.dhead"] if with_heads else None), (dyb_d if has_lm else None), 0
        denc = self.buf(a, "g.denc", (Me, d), F32)
        Hd, Fd = cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        pad_e = a["pad_e"] if a["has_mask_e"] else None
        first_cross = True
        for l in reversed(range(cfg.decoder_layers)):
            lp, tag = self.n(f"decoder.layers.{l}"), f"d{l}."
            self._ffn_block_bwd(bwd, a, tag, lp, dyA, dyB, Md, Fd, a[tag + "mid2_b16"], p_drop, 102 + 3 * l, acc, dpre_d[flip], dyb_d)
            dyA, dyB, flip = dpre_d[flip], dyb_d, flip ^ 1
            self._cross_block_bwd(bwd, a, tag, lp, dyA, dyB, Md, Me, B, Sd, Se, Hd, pad_e, a[tag + "mid_b16"], a["enc_b16"], p_drop,
                                  101 + 3 * l, acc, dpre_d[flip], dyb_d, denc, first_cross)
            first_cross = False
            dyA, dyB, flip = dpre_d[flip], dyb_d, flip ^ 1
            self._self_block_bwd(bwd, a, tag, lp, dyA, dyB, Md, B, Sd, Hd, a["pad_d"], True, a[tag + "in_b16"], p_drop, 100 + 3 * l, acc,
                                 dpre_d[flip], dyb_d)
            dyA, dyB, flip = dpre_d[flip], dyb_d, flip ^ 1
            if self.grad_reducer is not None:   # this layer's weight gradients are final: exchange them while the sweep goes on
                bwd.add(self.grad_reducer.launch_stage, cfg.decoder_layers - 1 - l)
        # decoder embedding
        demb_d = dpre_d[flip]
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        self.ln_bwd(bwd, dyA, dyB, a["d.emb_pre"], a["d.emb_mean"], a["d.emb_rstd"], self.n("decoder.layernorm_embedding"), demb_d, None, None, Md,
                    drop_in=(p_drop, 2))
        bwd.add(self.lib.kmb_embed_bwd, _ptr(demb_d), _ptr(a["ids_d"]), 0, _ptr(st.g(self.n("shared.weight"))), 0,
                _ptr(st.g(self.n("decoder.embed_positions.weight"))), B, Sd, d, cfg.extra_pos_embeddings, cfg.pad_token_id,
                scale, acc, bwd.stream)
        # encoder layers
        dpre_e = [self.buf(a, "g.dpreA_e", (Me, d), F32), self.buf(a, "g.dpreB_e", (Me, d), F32)]
        dyb_e = self.buf(a, "g.dyb_e", (Me, d), BF16)
        dyA, dyB, flip = denc, None, 0
        He, Fe = cfg.encoder_attention_heads, cfg.encoder_ffn_dim
        for l in reversed(range(cfg.encoder_layers)):
            lp, tag = self.n(f"encoder.layers.{l}"), f"e{l}."
            self._ffn_block_bwd(bwd, a, tag, lp, dyA, dyB, Me, Fe, a[tag + "x1_b16"], p_drop, 11 + 2 * l, acc, dpre_e[flip], dyb_e)
            dyA, dyB, flip = dpre_e[flip], dyb_e, flip ^ 1
            self._self_block_bwd(bwd, a, tag, lp, dyA, dyB, Me, B, Se, He, pad_e, False, a[tag + "in_b16"], p_drop, 10 + 2 * l, acc,
                                 dpre_e[flip], dyb_e)
            dyA, dyB, flip = dpre_e[flip], dyb_e, flip ^ 1
            if self.grad_reducer is not None:
                bwd.add(self.grad_reducer.launch_stage, cfg.decoder_layers + cfg.encoder_layers - 1 - l)
        demb_e = dpre_e[flip]
        dvis = self.buf(a, "g.dvis", (max(R, 1), d), BF16)
        if R > 0:
            bwd.add(_zero, dvis)     # rows no token points at (capacity tail, unused regions) must read as zero
        self.ln_bwd(bwd, dyA, dyB, a["e.emb_pre"], a["e.emb_mean"], a["e.emb_rstd"], self.n("encoder.layernorm_embedding"), demb_e, None, None, Me,
                    drop_in=(p_drop, 1))
        bwd.add(self.lib.kmb_embed_bwd, _ptr(demb_e), _ptr(a["ids_e"]), _ptr(a["slot"]), _ptr(st.g(self.n("shared.weight"))),
                _ptr(dvis), _ptr(st.g(self.n("encoder.embed_positions.weight"))), B, Se, d, cfg.extra_pos_embeddings,
                cfg.pad_token_id, scale, acc, bwd.stream)
        if R > 0:
            wname = self.n("encoder.embed_images.linear.weight")
            self.colsum(bwd, dvis, d, st.g(self.n("encoder.embed_images.linear.bias")), R, d)
            self.gemm(bwd, dvis, a["feats16"], d, self.fin - 4, R, d, self.fin - 4, a_mn=1, b_mn=1, out_f32=st.g(wname),
                      ld_f32=self.fin, accumulate=acc)
            bwd.add(self.lib.kmb_box_wgrad, _ptr(dvis), _ptr(a["boxes"]), _ptr(st.g(wname)), R, d, self.fin, bwd.stream)
        if self.grad_reducer is not None:
            bwd.add(self.grad_reducer.finish)
        return bwd

    # ------------------------------------------------------------------ input staging
    def _stage_inputs(self, a, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels):
        cfg = self.cfg
        a["ids_e"].copy_(input_ids.reshape(-1))
        if a["has_mask_e"]:
            a["amask_e"].copy_(attention_mask.reshape(-1))
        if decoder_input_ids is not None:
            a["ids_d"].copy_(decoder_input_ids.reshape(-1))
            if decoder_attention_mask is not None:
                torch.eq(decoder_attention_mask.reshape(-1), 0, out=a["pad_d_bool"])
            else:  # make_padding_mask(decoder_input_ids, pad) of HF-3.0.2 _prepare_bart_decoder_inputs
                torch.eq(a["ids_d"], cfg.pad_token_id, out=a["pad_d_bool"])
        if labels is not None:
            a["labels"].copy_(labels.reshape(-1))
        # pinned staging is double-buffered and guarded by events: a host that runs ahead (no loss.item(), the DeviceFeeder
        # path) must not overwrite a buffer whose host-to-device copy of the previous step has not executed yet
        i = a["stage_i"] = a.get("stage_i", 0) ^ 1
        ev = a["stage_ev"][i]
        if ev is not None:
            ev.synchronize()
        if isinstance(image_features, torch.Tensor):   # packed [R, 2052] fast path
            a["packed_buf"][:image_features.shape[0]].copy_(image_features)
        else:
            ptrs = [f.data_ptr() for f in image_features]
            a["feat_ptrs_host"][i].copy_(torch.tensor(ptrs, dtype=torch.int64))
            a["feat_ptrs"].copy_(a["feat_ptrs_host"][i], non_blocking=True)
        a["row_off_host"][i].copy_(torch.tensor(a["row_off_list"], dtype=torch.int32))
        a["row_off"].copy_(a["row_off_host"][i], non_blocking=True)
        if ev is None:
            ev = a["stage_ev"][i] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))

    def _get_arena(self, mode, input_ids, image_features, attention_mask, decoder_input_ids, labels, training,
                   lm_factor=1.0, image_counts=None, has_lm=True, with_heads=False):
        B, Se = input_ids.shape
        Sd = decoder_input_ids.shape[1] if decoder_input_ids is not None else 0
        packed = isinstance(image_features, torch.Tensor)
        if packed:
            assert image_counts is not None, "packed image_features need image_counts"
            counts = list(image_counts)
            assert image_features.dtype == F32 and image_features.is_contiguous()
        else:
            counts = [int(f.shape[0]) for f in image_features]
            for f in image_features:
                if f.numel():
                    assert f.is_cuda and f.dtype == F32 and f.is_contiguous() and f.shape[1] == self.fin, \
                        "image_features must be CUDA fp32 [n_i, 2052] tensors"
        R = sum(counts)
        stream = self.stream()
        # R (regions in the batch: 10-50 per image in the reference's data) is a CAPACITY of the workspace, not part of its
        # identity: rows [R, capacity) of the packed features / boxes / visual gradients are kept at zero, which every
        # consumer (GEMM rows, K-reductions, column sums) ignores.  A batch with more regions grows the workspace.
        key = (mode, B, Se, Sd, R > 0, attention_mask is not None, bool(training), packed, float(lm_factor), stream,
               bool(has_lm), bool(with_heads))
        a = self.arenas.get(key)
        grown = 0
        if a is not None and a["R"] < R:
            grown = a["R"] * 5 // 4
            self.arenas.pop(key)
            self.plans.pop(key, None)
            a = None
        if a is None:
            dev = self.device
            cap = (max(R, grown) + 63) // 64 * 64
            a = {"__dev": dev, "B": B, "Se": Se, "Sd": Sd, "R": cap, "training": bool(training),
                 "has_mask_e": attention_mask is not None, "stream": stream, "lm_factor": float(lm_factor),
                 "has_lm": bool(has_lm), "with_heads": bool(with_heads)}
            a["ids_e"] = torch.empty(B * Se, dtype=torch.int64, device=dev)
            a["amask_e"] = torch.empty(B * Se, dtype=torch.int64, device=dev)
            a["pad_e"] = torch.zeros(B * Se, dtype=torch.uint8, device=dev)
            a["ids_d"] = torch.empty(max(B * Sd, 1), dtype=torch.int64, device=dev)
            a["pad_d_bool"] = torch.zeros(max(B * Sd, 1), dtype=torch.bool, device=dev)
            a["pad_d"] = a["pad_d_bool"].view(torch.uint8)
            a["labels"] = torch.full((max(B * Sd, 1),), -100, dtype=torch.int64, device=dev)
            a["feat_ptrs"] = torch.zeros(B, dtype=torch.int64, device=dev)
            a["feat_ptrs_host"] = [torch.zeros(B, dtype=torch.int64).pin_memory() for _ in range(2)]
            a["row_off"] = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            a["row_off_host"] = [torch.zeros(B + 1, dtype=torch.int32).pin_memory() for _ in range(2)]
            a["stage_ev"] = [None, None]
            a["packed_buf"] = torch.zeros(max(cap, 1), self.fin, dtype=F32, device=dev) if packed else None
            a["packed"] = a["packed_buf"]
            a["loss"] = torch.zeros(1, dtype=F32, device=dev)
            a["seed"] = torch.zeros(1, dtype=torch.int64, device=dev)
            a["serial"] = 0
            a["flb"] = None
            self.remember(key, a)
        else:
            self.touch(key)
        off = [0]
        for c in counts:
            off.append(off[-1] + c)
        a["row_off_list"] = off
        return a, key

    # ------------------------------------------------------------------ public: training step pieces
    def train_forward(self, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels,
                      final_logits_bias, training, lm_factor=1.0, image_counts=None, with_heads=False):
        image_features, image_counts = unwrap_features(image_features, image_counts)   # kmbart.feed.PackedImageFeatures
        self.sync_shadow()
        a, key = self._get_arena("train", input_ids, image_features, attention_mask, decoder_input_ids, labels, training,
                                 lm_factor, image_counts, has_lm=labels is not None, with_heads=with_heads)
        a["flb"] = final_logits_bias.reshape(-1)
        self._stage_inputs(a, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels)
        plans = self.plans.get(key)
        if plans is None or plans["flb"] != a["flb"].data_ptr():
            plans = {"fwd": self._build_train_fwd(a), "flb": a["flb"].data_ptr()}
            self.plans[key] = plans
        a["heads_ctx"] = None
        self._fwd_serial += 1
        a["serial"] = self._fwd_serial       # stamps the activation stash: a later forward on this workspace overwrites it
        a["__plans"] = plans
        plans["fwd"].run()
        self.last_train = (a, key)
        self.launches_last = plans["fwd"].kernel_count()
        return a

    def grads_alias_flat_buffer(self):
        """True when existing .grad tensors still live in the flat gradient buffer (gradient
        accumulation, or zero_grad(set_to_none=False)): the backward must then add in place."""
        base, end = self.store.G.data_ptr(), self.store.G.data_ptr() + 4 * self.store.total
        return any(p.grad is not None and base <= p.grad.data_ptr() < end for p in self.store.params.values())

    def train_backward(self, a, key, upstream, accumulate, serial=None):
        if serial is not None and a["serial"] != serial:
            raise RuntimeError(
                "backward through a KM-BART training forward whose activation stash has been overwritten: another training-mode "
                "forward with the same input shape ran on this model before this loss was back-propagated (the fused step keeps "
                "ONE stash per shape).  Call backward() before the next forward of the same shape, or sum the losses of "
                "different-shape micro-batches instead.")
        if upstream is not None:
            self.upstream.copy_(upstream.reshape(1).to(F32))
        else:
            self.upstream.fill_(1.0)
        plans = a["__plans"]       # survives an LRU eviction of the workspace between forward and backward
        if a.get("heads_ctx"):
            from .heads import heads_backward
            heads_backward(self, a, a["heads_ctx"], accumulate)
        name = "bwd1" if accumulate else "bwd0"
        if name not in plans:
            plans[name] = self._build_train_bwd(a, accumulate)
        plans[name].run()
        self.launches_last += plans[name].kernel_count()

    # ------------------------------------------------------------------ public: inference forward (no cache)
    def infer_forward(self, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask,
                      encoder_only=False, image_counts=None):
        """Full-sequence forward without stashing; returns (enc_f32 [B,Se,d], dec_f32 [B,Sd,d] or None, arena)."""
        image_features, image_counts = unwrap_features(image_features, image_counts)
        self.sync_shadow()
        a, key = self._get_arena("enc" if encoder_only else "infer", input_ids, image_features, attention_mask,
                                 None if encoder_only else decoder_input_ids, None, False, 1.0, image_counts)
        self._stage_inputs(a, input_ids, image_features, attention_mask, None if encoder_only else decoder_input_ids,
                           decoder_attention_mask, None)
        plan = self.plans.get(key)
        if plan is None:
            plan = Plan()
            plan.stream = a["stream"]
            B, Se, Sd, R = a["B"], a["Se"], a["Sd"], a["R"]
            _, enc_b16 = self._encoder_fwd(plan, a, B, Se, R, False, 0.0)
            if not encoder_only:
                self._decoder_fwd(plan, a, B, Sd, Se, False, 0.0, enc_b16)
            self.plans[key] = plan
        plan.run()
        self.launches_last = len(plan)
        d = self.cfg.d_model
        enc = a["enc_f32"].view(a["B"], a["Se"], d)
        dec = None if encoder_only else a["dec_f32"].view(a["B"], a["Sd"], d)
        return enc, dec, a

    def logits_from_hidden(self, h_b16, final_logits_bias, out=None):
        """Materialise logits = h E^T + final_logits_bias (fp32) — only on demand (generation's last
        position, scripts/filter_reason.py:42, pretrain.py:285); the training loss never does."""
        st, d, V = self.store, self.cfg.d_model, self.cfg.vocab_size
        M = h_b16.shape[0]
        if out is None:
            out = torch.empty(M, V, dtype=F32, device=self.device)
        plan = Plan()
        plan.stream = self.stream()
        self.gemm(plan, h_b16, st.p16(self.n("shared.weight")), M, V, d, d, d, bias=final_logits_bias.reshape(-1), out_f32=out, ld_f32=V)
        plan.run()
        return out

    # ------------------------------------------------------------------ decoder-only passes (generation / sample_sentence)
    def decoder_full(self, enc_hidden, attention_mask, decoder_input_ids, decoder_attention_mask):
        """Decoder over a whole target prefix given encoder states [B, Se, d] (use_cache=False path of
        src/model/model.py:39-103 with encoder_outputs supplied, as src/model/utils.py:19-28 calls it)."""
        self.sync_shadow()
        cfg, d = self.cfg, self.cfg.d_model
        B, Se = enc_hidden.shape[0], enc_hidden.shape[1]
        Sd = decoder_input_ids.shape[1]
        stream = self.stream()
        key = ("decfull", B, Se, Sd, attention_mask is not None, stream)
        a = self.arenas.get(key)
        if a is None:
            dev = self.device
            a = {"__dev": dev, "B": B, "Se": Se, "Sd": Sd, "R": 0, "has_mask_e": attention_mask is not None, "stream": stream}
            a["amask_e"] = torch.empty(B * Se, dtype=torch.int64, device=dev)
            a["pad_e"] = torch.zeros(B * Se, dtype=torch.uint8, device=dev)
            a["ids_d"] = torch.empty(B * Sd, dtype=torch.int64, device=dev)
            a["pad_d_bool"] = torch.zeros(B * Sd, dtype=torch.bool, device=dev)
            a["pad_d"] = a["pad_d_bool"].view(torch.uint8)
            a["enc_in_b16"] = torch.empty(B * Se, d, dtype=BF16, device=dev)
            self.arenas[key] = a
        a["enc_in_b16"].copy_(enc_hidden.reshape(B * Se, d))
        a["ids_d"].copy_(decoder_input_ids.reshape(-1))
        if decoder_attention_mask is not None:
            torch.eq(decoder_attention_mask.reshape(-1), 0, out=a["pad_d_bool"])
        else:
            torch.eq(a["ids_d"], cfg.pad_token_id, out=a["pad_d_bool"])
        if attention_mask is not None:
            a["amask_e"].copy_(attention_mask.reshape(-1))
        plan = self.plans.get(key)
        if plan is None:
            plan = Plan()
            plan.stream = stream
            if attention_mask is not None:
                plan.add(self.lib.kmb_invert_mask, _ptr(a["amask_e"]), _ptr(a["pad_e"]), B * Se, stream)
            self._decoder_fwd(plan, a, B, Sd, Se, False, 0.0, a["enc_in_b16"])
            self.plans[key] = plan
        plan.run()
        self.launches_last = len(plan)
        return a["dec_f32"].view(B, Sd, d), a

    def _attn_strided(self, plan, q4, k4, v4, o4, pad, n, H, Sq, Sk):
        """q4/k4/v4/o4: logical [n, H, S, 64] bf16 tensors with arbitrary (multiple-of-8) strides."""
        arr = (C.c_int64 * 12)()
        for i, t in enumerate((q4, k4, v4, o4)):
            assert t.stride(3) == 1
            arr[3 * i], arr[3 * i + 1], arr[3 * i + 2] = t.stride(0), t.stride(1), t.stride(2) if t.shape[2] > 1 else 64
        plan.add(self.lib.kmb_attn_fwd_strided, _ptr(q4), _ptr(k4), _ptr(v4), _ptr(o4), arr, _ptr(pad), n, H, Sq, Sk, 64, 0,
                 0.125, plan.stream, keep=arr)

    def decoder_step(self, last_ids, position, enc_hidden, enc_pad_u8, caches):
        """One cached decode step (HF-3.0.2 BartDecoder.forward with use_cache=True): embeds the last
        token at `position`, appends its K/V to the per-layer self-attention cache, reuses (or on the
        first step builds) the static cross-attention K/V.  Cache tensors keep the legacy logical
        shape [rows, heads, T, 64] so src/model/mixins.py:419-434 style re-ordering works on them.
        Returns (hidden bf16 [rows, d], new_caches)."""
        self.sync_shadow()
        cfg, st, d = self.cfg, self.store, self.cfg.d_model
        n = last_ids.shape[0]
        H, F = cfg.decoder_attention_heads, cfg.decoder_ffn_dim
        dev = self.device
        plan = Plan()
        plan.stream = self.stream()
        ids = last_ids.reshape(-1).contiguous()
        x_f32 = torch.empty(n, d, dtype=F32, device=dev)
        x_b16 = torch.empty(n, d, dtype=BF16, device=dev)
        scale = math.sqrt(d) if cfg.scale_embedding else 1.0
        plan.add(self.lib.kmb_embed_ln_fwd, _ptr(ids), 0, _ptr(st.p32(self.n("shared.weight"))),
                 _ptr(st.p32(self.n("decoder.embed_positions.weight"))), 0, 0, 0, 0,
                 _ptr(st.p32(self.n("decoder.layernorm_embedding.weight"))),
                 _ptr(st.p32(self.n("decoder.layernorm_embedding.bias"))), 0, _ptr(x_f32), _ptr(x_b16), 0, 0, n, 1, d,
                 cfg.extra_pos_embeddings + int(position), 0, scale, 0.0, 0, 0, plan.stream)
        keep = [ids, x_f32, x_b16]
        new_caches = []
        enc_b16 = None
        for l in range(cfg.decoder_layers):
            lp = self.n(f"decoder.layers.{l}")
            lc = caches[l] if caches is not None else {}
            # --- self attention with growing cache
            qkv = torch.empty(n, 3 * d, dtype=BF16, device=dev)
            self.gemm(plan, x_b16, st.p16(lp + ".self_attn.q_proj.weight", 3 * d), n, 3 * d, d, d, d,
                      bias=st.fused32(lp + ".self_attn.q_proj.bias", 3), out_bf16=qkv)
            plan.run(); plan.calls.clear()
            k_new = qkv[:, d:2 * d].view(n, 1, H, 64).permute(0, 2, 1, 3)
            v_new = qkv[:, 2 * d:].view(n, 1, H, 64).permute(0, 2, 1, 3)
            sc = lc.get("self")
            if sc is not None and sc.get("prev_key") is not None:
                K = torch.cat([sc["prev_key"], k_new], dim=2)
                V = torch.cat([sc["prev_value"], v_new], dim=2)
            else:
                K, V = k_new.contiguous(), v_new.contiguous()
            T = K.shape[2]
            ctx = torch.empty(n, d, dtype=BF16, device=dev)
            self._attn_strided(plan, qkv[:, :d].view(n, 1, H, 64).permute(0, 2, 1, 3), K, V,
                               ctx.view(n, 1, H, 64).permute(0, 2, 1, 3), None, n, H, 1, T)
            pre = torch.empty(n, d, dtype=BF16, device=dev)
            y_f32 = torch.empty(n, d, dtype=F32, device=dev)
            y_b16 = torch.empty(n, d, dtype=BF16, device=dev)
            self.gemm(plan, ctx, st.p16(lp + ".self_attn.out_proj.weight"), n, d, d, d, d,
                      bias=st.p32(lp + ".self_attn.out_proj.bias"), out_bf16=pre)
            self.ln_fwd(plan, pre, x_f32, lp + ".self_attn_layer_norm", None, y_f32, y_b16, None, None, n)
            # --- cross attention with static cache
            cc = lc.get("encoder_decoder")
            q2 = torch.empty(n, d, dtype=BF16, device=dev)
            self.gemm(plan, y_b16, st.p16(lp + ".encoder_attn.q_proj.weight"), n, d, d, d, d,
                      bias=st.p32(lp + ".encoder_attn.q_proj.bias"), out_bf16=q2)
            if cc is not None and cc.get("prev_key") is not None:
                K2, V2 = cc["prev_key"], cc["prev_value"]
            else:
                Se = enc_hidden.shape[1]
                if enc_b16 is None:
                    enc_b16 = enc_hidden.reshape(n * Se, d).to(BF16)
                kv2 = torch.empty(n * Se, 2 * d, dtype=BF16, device=dev)
                self.gemm(plan, enc_b16, st.p16(lp + ".encoder_attn.k_proj.weight", 2 * d), n * Se, 2 * d, d, d, d,
                          bias=st.fused32(lp + ".encoder_attn.k_proj.bias", 2), out_bf16=kv2)
                kv5 = kv2.view(n, Se, 2, H, 64)
                K2, V2 = kv5[:, :, 0].permute(0, 2, 1, 3), kv5[:, :, 1].permute(0, 2, 1, 3)
            Se = K2.shape[2]
            ctx2 = torch.empty(n, d, dtype=BF16, device=dev)
            self._attn_strided(plan, q2.view(n, 1, H, 64).permute(0, 2, 1, 3), K2, V2,
                               ctx2.view(n, 1, H, 64).permute(0, 2, 1, 3), enc_pad_u8, n, H, 1, Se)
            pre2 = torch.empty(n, d, dtype=BF16, device=dev)
            z_f32 = torch.empty(n, d, dtype=F32, device=dev)
            z_b16 = torch.empty(n, d, dtype=BF16, device=dev)
            self.gemm(plan, ctx2, st.p16(lp + ".encoder_attn.out_proj.weight"), n, d, d, d, d,
                      bias=st.p32(lp + ".encoder_attn.out_proj.bias"), out_bf16=pre2)
            self.ln_fwd(plan, pre2, y_f32, lp + ".encoder_attn_layer_norm", None, z_f32, z_b16, None, None, n)
            # --- FFN
            hbuf = torch.empty(n, F, dtype=BF16, device=dev)
            pre3 = torch.empty(n, d, dtype=BF16, device=dev)
            x_f32 = torch.empty(n, d, dtype=F32, device=dev)
            x_b16 = torch.empty(n, d, dtype=BF16, device=dev)
            self.gemm(plan, z_b16, st.p16(lp + ".fc1.weight"), n, F, d, d, d, bias=st.p32(lp + ".fc1.bias"), act=L.ACT_GELU, out_bf16=hbuf)
            self.gemm(plan, hbuf, st.p16(lp + ".fc2.weight"), n, d, F, F, F, bias=st.p32(lp + ".fc2.bias"), out_bf16=pre3)
            self.ln_fwd(plan, pre3, z_f32, lp + ".final_layer_norm", None, x_f32, x_b16, None, None, n)
            plan.run(); plan.calls.clear()
            keep += [qkv, ctx, pre, y_f32, y_b16, q2, ctx2, pre2, z_f32, z_b16, hbuf, pre3, K, V, K2, V2]
            new_caches.append({"self": {"prev_key": K, "prev_value": V, "prev_key_padding_mask": None},
                               "encoder_decoder": {"prev_key": K2, "prev_value": V2, "prev_key_padding_mask": None}})
        return x_b16, x_f32, new_caches
