"""Data-parallel gradient exchange for the fused training step.

The reference wraps the model in torch DDP (`DDP(model, find_unused_parameters=True)`, vcg_train.py:96-98):
bucketed NCCL all-reduce of the gradients driven by autograd hooks, overlapped with backward.  That still works
with this implementation, but the fused backward is ONE autograd node, so DDP's hooks only fire
after the last kernel and nothing overlaps.  FlatGradReducer uses what the engine already has — every gradient
lives in one flat fp32 buffer and the backward is an ordered launch plan — to do better with less machinery:

  * the flat buffer is cut into contiguous regions by the point of the backward sweep at which they are complete
    (decoder layer L-1 ... 0, encoder layer L-1 ... 0, then "the rest": small tensors, image projection,
    positions and the tied embedding, whose gradient is only final after the encoder embedding backward);
  * the plan launches `all_reduce(region, AVG, async_op=True)` right after the kernel that completes a region —
    NCCL runs it on its own stream over NVLink while the sweep continues; no bucket copies, no hooks;
  * `finish()` (last plan entry) joins the NCCL stream.

One process per GPU, launched by torchrun; works on any torch.distributed backend (tests use gloo on CPU)."""
import torch
import torch.distributed as dist


def layer_stage(name, n_dec, n_enc):
    """Backward-order stage at which the gradient of parameter `name` is complete (None = end of the sweep)."""
    parts = name.split(".")
    if "layers" in parts:
        i = parts.index("layers")
        layer = int(parts[i + 1])
        if parts[i - 1] == "decoder":
            return n_dec - 1 - layer
        if parts[i - 1] == "encoder":
            return n_dec + (n_enc - 1 - layer)
    return None


def plan_regions(names, offsets, numels, small_end, total, n_dec, n_enc):
    """[(start, end, stage)] covering [0, total) exactly once; stage None = reduced at the end."""
    items = sorted((offsets[n], numels[n], n) for n in names)
    regions = []
    if small_end > 0:
        regions.append([0, small_end, None])
    for off, num, n in items:
        if off < small_end:
            continue
        st = layer_stage(n, n_dec, n_enc)
        if regions and regions[-1][2] == st:
            regions[-1][1] = off + num
        else:
            start = regions[-1][1] if regions else 0   # alignment gaps travel with the following region
            regions.append([start, off + num, st])
    if regions:
        regions[-1][1] = total
    else:
        regions.append([0, total, None])
    return [tuple(r) for r in regions]


class FlatGradReducer:
    def __init__(self, model, process_group=None, broadcast_parameters=True, engine=None, defer_tail=False):
        """`engine`: anything with `.store` (a kmbart.engine.ParamStore), `.cfg`, `.plans` — defaults to the model's
        sm_100a engine; the CPU/gloo tests pass a stand-in that only owns a ParamStore.
        `defer_tail=True`: finish() does not join the exchange of the LAST regions (tied embedding + small tensors:
        their gradient is only final when the sweep ends, so this all-reduce can never overlap with backward);
        `kmbart.optim.AdamW.step()` then updates every other parameter first and joins (`wait_tail()`) before it
        touches the deferred ranges.  Anything else that reads `.grad` between backward and the optimizer step
        (clipping, GradScaler.unscale_) must call `wait_tail()` itself — hence opt-in."""
        self.model = model
        self.defer_tail = bool(defer_tail)
        self.tail_pending = []
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        eng = engine if engine is not None else model._engine()
        self.eng = eng
        st = eng.store
        cfg = eng.cfg
        numels = {n: p.numel() for n, p in st.params.items()}
        self.regions = plan_regions(st.names, st.offsets, numels, st.small_end, st.total, cfg.decoder_layers, cfg.encoder_layers)
        self.enabled = True
        self.pending = []
        self.bytes_reduced = 0
        self.tail_ranges = [(a, b) for a, b, s_ in self.regions if s_ is None]   # element ranges of the flat buffers
        st.grad_reducer = self
        eng.grad_reducer = self
        eng.plans = {k: ({kk: vv for kk, vv in v.items() if not kk.startswith("bwd")} if isinstance(v, dict) else v)
                     for k, v in eng.plans.items()}   # backward plans are rebuilt with the reduction points
        if broadcast_parameters and self.world > 1:
            dist.broadcast(st.P, src=0, group=process_group)     # DDP constructor semantics: rank 0's weights everywhere
            for b in model.buffers():
                dist.broadcast(b, src=0, group=process_group)
            st.shadow_version = None

    # ---- called from the backward launch plan
    def launch_stage(self, stage):
        if not self.enabled or self.world == 1:
            return 0
        G = self.eng.store.G
        for a, b, s in self.regions:
            if s == stage:
                self.pending.append(dist.all_reduce(G[a:b], op=dist.ReduceOp.AVG if G.is_cuda else dist.ReduceOp.SUM,
                                                    group=self.group, async_op=True))
                if not G.is_cuda:
                    self._scale_cpu = True
                self.bytes_reduced += 4 * (b - a)
        return 0

    def finish(self):
        n_before = len(self.pending)
        self.launch_stage(None)
        tail = self.pending[n_before:]
        head = self.pending[:n_before]
        self.pending = []
        for w in head:
            w.wait()
        if self.defer_tail and self.eng.store.G.is_cuda:
            self.tail_pending = tail          # joined by AdamW.step() / wait_tail()
        else:
            for w in tail:
                w.wait()
        if getattr(self, "_scale_cpu", False):   # gloo has no AVG
            self.eng.store.G.div_(self.world)
            self._scale_cpu = False
        return 0

    def wait_tail(self):
        """Orders the current stream after the deferred exchange of the last regions (no-op when nothing is pending)."""
        for w in self.tail_pending:
            w.wait()
        self.tail_pending = []

    class _NoSync:
        def __init__(self, r):
            self.r = r

        def __enter__(self):
            self.prev, self.r.enabled = self.r.enabled, False

        def __exit__(self, *a):
            self.r.enabled = self.prev

    def no_sync(self):
        """Gradient accumulation: skip the exchange for micro-steps inside this context (DDP.no_sync semantics)."""
        return FlatGradReducer._NoSync(self)
