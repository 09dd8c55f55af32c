"""Multitask pre-training forward/backward — the host side of
MultiModalBartForPreTraining.forward (reference src/model/model.py:162-309).

The trunk (encoder, decoder, fused LM-head + cross-entropy with `lm_loss_factor` and the
`labels == cls_token_id -> -100` rewrite, :296-301) runs through the same launch plans as
fine-tuning.  The three BartClassificationHead losses are evaluated on gathered decoder states:
    MRM        KL-div(log_softmax(head(h[mrm_mask])), soft labels), batchmean      (:246-257)
    attribute  CE(head(h[attribute_mask]), labels)                                 (:259-268)
    relation   CE(head(cat(h[b, obj], h[b, subj])), labels)                        (:270-289)
with kernels  gather_rows_bf16 -> tcgen05 GEMM (tanh epilogue) -> GEMM (fp32 logits) ->
small_xent,  and in backward  small_xent(dlogits) -> wgrad / dgrad GEMMs (tanh' epilogue) ->
scatter_add_rows into the fp32 gradient of the decoder output that the trunk backward consumes.
The reference's boolean-mask gathers force a device sync per head; here the row indices come from
a stable device sort of the mask (row counts are known from the label lists), and relation pairs
(host dicts in the reference API) are flattened once on the host."""
import torch

from . import lib as L
from .engine import Plan, BF16, F32, _ptr


def _rup(x, m):
    return (x + m - 1) // m * m


class _Head:
    def __init__(self, name, din, classes, factor, mode):
        self.name, self.din, self.classes, self.factor, self.mode = name, din, classes, float(factor), mode


def _mask_rows(mask, n):
    """indices (int32) of the first n True entries of the flattened mask, in order, without a host sync"""
    flat = mask.reshape(-1).to(torch.bool)
    order = torch.sort((~flat).to(torch.uint8), stable=True).indices
    return order[:n].to(torch.int32).contiguous()


def _head_forward(eng, a, plan, head, idx_list, n, labels, soft, hstate):
    """idx_list: one int32 index tensor per din/d block.  Returns the per-head context kept for backward."""
    st, d, lib = eng.store, eng.cfg.d_model, eng.lib
    dev = eng.device
    C_ = head.classes
    ldf, ldd = _rup(C_, 4), _rup(C_, 8)
    rep = torch.empty(n, head.din, dtype=BF16, device=dev)
    for j, idx in enumerate(idx_list):
        plan.add(lib.kmb_gather_rows_bf16, _ptr(a["dec_b16"]), d, _ptr(idx), rep.data_ptr() + 2 * j * d, head.din, n, d, plan.stream)
    t = torch.empty(n, d, dtype=BF16, device=dev)
    logits = torch.zeros(n, ldf, dtype=F32, device=dev)
    loss = torch.zeros(1, dtype=F32, device=dev)
    eng.gemm(plan, rep, st.p16(head.name + ".dense.weight"), n, d, head.din, head.din, head.din,
             bias=st.p32(head.name + ".dense.bias"), act=L.ACT_TANH, out_bf16=t)
    eng.gemm(plan, t, st.p16(head.name + ".out_proj.weight"), n, C_, d, d, d, bias=st.p32(head.name + ".out_proj.bias"),
             out_f32=logits, ld_f32=ldf)
    plan.add(lib.kmb_small_xent, _ptr(logits), ldf, n, C_, head.mode, _ptr(labels), _ptr(soft), C_, head.factor, _ptr(loss), 0, 0, 0,
             plan.stream)
    return dict(head=head, idx=idx_list, n=n, rep=rep, t=t, logits=logits, loss=loss, labels=labels, soft=soft, ldf=ldf, ldd=ldd)


def _head_backward(eng, a, plan, hc):
    st, d, lib = eng.store, eng.cfg.d_model, eng.lib
    head, n, C_ = hc["head"], hc["n"], hc["head"].classes
    dev = eng.device
    ldd = hc["ldd"]
    dlog = torch.empty(n, ldd, dtype=BF16, device=dev)
    dpre = torch.empty(n, d, dtype=BF16, device=dev)
    drep = torch.empty(n, head.din, dtype=BF16, device=dev)
    dbo = torch.zeros(ldd, dtype=F32, device=dev)
    plan.add(lib.kmb_small_xent, _ptr(hc["logits"]), hc["ldf"], n, C_, head.mode, _ptr(hc["labels"]), _ptr(hc["soft"]), C_, head.factor,
             0, _ptr(dlog), ldd, _ptr(eng.upstream), plan.stream)
    Wo, Wd = head.name + ".out_proj.weight", head.name + ".dense.weight"
    # dWo[C, d] += dlog^T t ; dbo = colsum(dlog) ; dpre = (dlog Wo) * (1 - t^2)
    eng.gemm(plan, dlog, hc["t"], C_, d, n, ldd, d, a_mn=1, b_mn=1, out_f32=st.g(Wo), ld_f32=d, accumulate=1)
    eng.colsum(plan, dlog, ldd, dbo, n, ldd)
    eng.gemm(plan, dlog, st.p16(Wo), n, d, C_, ldd, d, b_mn=1, act=L.ACT_TANH_GRAD, aux=hc["t"], ld_aux=d, out_bf16=dpre)
    # dWd[d, din] += dpre^T rep ; dbd = colsum(dpre) ; drep = dpre Wd
    eng.gemm(plan, dpre, hc["rep"], d, head.din, n, d, head.din, a_mn=1, b_mn=1, out_f32=st.g(Wd), ld_f32=head.din, accumulate=1)
    eng.colsum(plan, dpre, d, st.g(head.name + ".dense.bias"), n, d)
    eng.gemm(plan, dpre, st.p16(Wd), n, head.din, d, d, head.din, b_mn=1, out_bf16=drep)
    for j, idx in enumerate(hc["idx"]):
        plan.add(lib.kmb_scatter_add_rows, drep.data_ptr() + 2 * j * d, head.din, _ptr(idx), _ptr(a["g.dhead"]), d, n, d, plan.stream)
    hc["bwd_keep"] = (dlog, dpre, drep, dbo)
    return dbo


def heads_backward(eng, a, ctx, accumulate):
    """Runs before the trunk's backward plan: fills the head parameter gradients and a['g.dhead']."""
    st = eng.store
    if not accumulate:
        st.G[:st.zero_end].zero_()
    plan = Plan()
    plan.stream = eng.stream()
    pending = []
    for hc in ctx["heads"]:
        pending.append((hc, _head_backward(eng, a, plan, hc)))
    plan.run()
    for hc, dbo in pending:   # class counts (1601 / 129) are odd: the column-sum kernel works on the padded width
        st.g(hc["head"].name + ".out_proj.bias").add_(dbo[:hc["head"].classes])
    eng.launches_last += len(plan) + len(pending)


class _PretrainStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, arena, key, *params):
        ctx.owner, ctx.arena, ctx.key, ctx.n_params, ctx.serial = owner, arena, key, len(params), arena["serial"]
        hc = arena.get("heads_ctx")
        # heads that took no part in this batch get grad None, like plain autograd in the reference (HF AdamW then skips
        # them: no momentum step, no step-count increment) instead of a zero gradient
        ctx.active_heads = {x["loss_name"].replace("_loss", "_head") for x in hc["heads"]} if hc else set()
        return arena["loss"].clone().squeeze(0)

    @staticmethod
    def backward(ctx, grad_loss):
        owner = ctx.owner
        eng = owner._engine()
        accumulate = eng.grads_alias_flat_buffer()
        eng.train_backward(ctx.arena, ctx.key, grad_loss, accumulate, serial=ctx.serial)
        if accumulate:
            return (None, None, None) + (None,) * ctx.n_params
        store = eng.store

        def inactive(n):
            top = n.split(".", 1)[0]
            return top.endswith("_head") and top not in ctx.active_heads
        return (None, None, None) + tuple(store.grad_view(n) if (p.requires_grad and not inactive(n)) else None
                                          for n, p in owner._named_params_cache)


def pretraining_forward(model, input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, labels,
                        mrm_labels, mrm_mask, attribute_labels, attribute_mask, relation_labels):
    from src.model.model import LazyLogits, _shift_tokens_right
    cfg = model.config
    eng = model._engine()
    dev = eng.device
    if decoder_input_ids is None:
        decoder_input_ids = _shift_tokens_right(input_ids, cfg.pad_token_id)
    lm_labels = None
    if labels is not None:
        lm_labels = labels.clone()
        lm_labels[lm_labels == model.cls_token_id] = -100      # src/model/model.py:297-298
    # ---- which heads are active (the reference skips a head whose gathered set is empty)
    specs = []
    if mrm_labels is not None:
        n = sum(int(x.shape[0]) for x in mrm_labels)
        if n > 0:
            soft = torch.cat([x.to(dev, F32) for x in mrm_labels], 0).contiguous()
            specs.append((_Head("mrm_head", cfg.d_model, cfg.num_labels, cfg.mrm_loss_factor, 1), [_mask_rows(mrm_mask, n)], n,
                          None, soft, "mrm_loss"))
    if attribute_labels is not None:
        n = sum(int(x.numel()) for x in attribute_labels)
        if n > 0:
            lab = torch.cat([x.reshape(-1) for x in attribute_labels], 0).to(dev, torch.int64).contiguous()
            specs.append((_Head("attribute_head", cfg.d_model, cfg.num_attributes, cfg.attribute_loss_factor, 0),
                          [_mask_rows(attribute_mask, n)], n, lab, None, "attribute_loss"))
    if relation_labels is not None:
        Sd = decoder_input_ids.shape[1]
        obj, subj, lab = [], [], []
        for b, rels in enumerate(relation_labels):
            for r in rels:
                obj.append(b * Sd + int(r["object_index"]))
                subj.append(b * Sd + int(r["subject_index"]))
                lab.append(int(r["label"]))
        if lab:
            packed = torch.tensor([obj, subj, lab], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
            specs.append((_Head("relation_head", 2 * cfg.d_model, cfg.num_relations, cfg.relation_loss_factor, 0),
                          [packed[0].to(torch.int32).contiguous(), packed[1].to(torch.int32).contiguous()], len(lab),
                          packed[2].contiguous(), None, "relation_loss"))
    with_heads = bool(specs)
    grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    a = eng.train_forward(input_ids, image_features, attention_mask, decoder_input_ids, decoder_attention_mask, lm_labels,
                          model.final_logits_bias, model.training, lm_factor=float(cfg.lm_loss_factor), with_heads=with_heads)
    key = eng.last_train[1]
    losses = {}
    ctx = {"heads": []}
    if with_heads:
        plan = Plan()
        plan.stream = eng.stream()
        for head, idx, n, lab, soft, lname in specs:
            hc = _head_forward(eng, a, plan, head, idx, n, lab, soft, None)
            hc["loss_name"] = lname
            ctx["heads"].append(hc)
        plan.run()
        eng.launches_last += len(plan)
        for hc in ctx["heads"]:
            a["loss"].add_(hc["loss"])
            losses[hc["loss_name"]] = hc["loss"].squeeze(0)
    a["heads_ctx"] = ctx if with_heads else None
    if lm_labels is not None:
        losses["lm_loss"] = a["lm_loss"].clone().squeeze(0)
    if grad:
        if not hasattr(model, "_named_params_cache"):
            model._named_params_cache = list(model.named_parameters())
        params = [p for _, p in model._named_params_cache]
        losses["loss"] = _PretrainStep.apply(model, a, key, *params)
    else:
        losses["loss"] = a["loss"].clone().squeeze(0)
    B, Sd, d = a["B"], a["Sd"], cfg.d_model
    h_copy = a["dec_b16"].clone()
    lazy = LazyLogits(lambda: model._logits(h_copy, B, Sd), (B, Sd, cfg.vocab_size), h_copy.device)
    enc = a["enc_f32"].view(B, a["Se"], d)
    return (losses, lazy, enc)
