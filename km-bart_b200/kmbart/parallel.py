"""Data-parallel gradient exchange for the fused training step.

The reference wraps the model in torch DDP (`DDP(model, find_unused_parameters=True)`, vcg_train.py:96-98):
bucketed NCCL all-reduce of the gradients driven by autograd hooks, overlapped with backward.  That still works
with this implementation, but the fused backward is ONE autograd node, so DDP's hooks only fire
after the last kernel and nothing overlaps.  FlatGradReducer uses what the engine already has — every gradient
lives in one flat fp32 buffer and the backward is an ordered launch plan — to do better with less machinery:

  * the flat buffer is cut into contiguous regions by the point of the backward sweep at which they are complete
    (decoder layer L-1 ... 0, encoder layer L-1 ... 0, then "the rest": small tensors, image projection,
    positions and the tied embedding, whose gradient is only final after the encoder embedding backward);
  * the plan starts the exchange of a region right after the kernel that completes it, on streams of its own,
    while the sweep continues; no bucket copies, no hooks;
  * `finish()` (last plan entry) joins; with `defer_tail=True` the regions that cannot finish before the sweep ends
    stay in flight and `kmbart.optim.AdamW.step()` joins them after it has updated everything else.

Transport: on one NVSwitch node (every rank a CUDA peer of every other) `PeerExchange` — csrc/peer_exchange.cu,
copy-engine pushes between CUDA-IPC mappings of the ranks' gradient buffers plus one small reduction kernel per
region, flags by stream memory operations: no SM is taken from the backward sweep (profiles/r02_dp_timeline.md).
Otherwise (or with KMBART_GRAD_EXCHANGE=nccl) `all_reduce(region, AVG, async_op=True)` per region; gloo on CPU (tests).

One process per GPU, launched by torchrun."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist

PEER_PIECE = 12 * 1024 * 1024   # elements (48 MB): larger regions are exchanged in pieces
_ipc_opened = {}     # handle bytes -> mapped base address (a process may map a peer allocation only once; never closed)


class PeerExchange:
    """Peer-memory gradient exchange for one flat gradient buffer (csrc/peer_exchange.cu).  `create` is collective:
    every rank of `group` calls it; it returns None on every rank when any pair of ranks is not a CUDA peer on one
    host (the caller then keeps NCCL)."""

    @staticmethod
    def create(G, regions, group=None):
        from . import lib as L
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 8 or not G.is_cuda:
            return None
        lib = L.load()
        dev = G.device.index if G.device.index is not None else torch.cuda.current_device()
        chunk = max((((b - a + world - 1) // world) + 3) // 4 * 4 for a, b, _ in regions)
        n_flags = 2 * len(regions) * world
        flag_words = (n_flags + 63) // 64 * 64
        # flags first (zeroed), then `world` staging slots; one allocation so one IPC mapping covers both
        side = torch.zeros(flag_words + 2 * world * chunk, dtype=torch.float32, device=G.device)   # two lanes of slots
        mine, ok = [], True
        try:
            for t in (G, side):
                h = (C.c_ubyte * 64)()
                off = C.c_ulonglong(0)
                L.check(lib.kmb_ipc_export(t.data_ptr(), h, C.byref(off)), "kmb_ipc_export")
                mine.append((bytes(h), int(off.value)))
        except L.KmbartError:
            ok = False
        info = [None] * world
        dist.all_gather_object(info, (socket.gethostname(), dev, ok, mine), group=group)
        ok = all(i[2] for i in info) and len({i[0] for i in info}) == 1 and len({i[1] for i in info}) == world
        ok = ok and all(lib.kmb_peer_can_access(dev, i[1]) for r, i in enumerate(info) if r != rank)
        peer_g, peer_side = [0] * world, [0] * world
        if ok:
            try:
                for r, i in enumerate(info):
                    if r == rank:
                        continue
                    ptrs = []
                    for hb, off in i[3]:
                        base = _ipc_opened.get(hb)
                        if base is None:
                            out = C.c_void_p()
                            L.check(lib.kmb_ipc_open(hb, C.byref(out)), "kmb_ipc_open")
                            base = _ipc_opened[hb] = out.value
                        ptrs.append(base + off)
                    peer_g[r], peer_side[r] = ptrs
            except L.KmbartError:
                ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            return None
        self = PeerExchange()
        self.lib, self.L, self.side, self.G = lib, L, side, G
        self.world, self.rank, self.regions = world, rank, list(regions)
        arr = lambda v: (C.c_void_p * world)(*[x or None for x in v])
        ctx = C.c_void_p()
        L.check(lib.kmb_peer_ctx_create(rank, world, G.data_ptr(), arr(peer_g), side.data_ptr() + 4 * flag_words,
                                        arr([p + 4 * flag_words if p else 0 for p in peer_side]), side.data_ptr(), arr(peer_side),
                                        chunk, len(regions), C.byref(ctx)), "kmb_peer_ctx_create")
        self.ctx = ctx
        self.value = 0
        tail = os.environ.get("KMBART_PEER_TAIL", "auto")      # ce | kernel | auto (kernel when world > 2)
        self.kernel_tail = tail == "kernel" or (tail == "auto" and world > 2)
        torch.cuda.synchronize(G.device)
        dist.barrier(group=group)          # nobody signals before every rank has zeroed and mapped its flags
        return self

    def exchange(self, region_index, after_sweep=False):
        """`after_sweep`: the region is exchanged when the backward sweep is over (same value on every rank) — with more
        than two ranks it then goes through the one-kernel load/store path instead of the copy engines."""
        a, b, _ = self.regions[region_index]
        stream = torch.cuda.current_stream(self.G.device).cuda_stream
        mode = 1 if (after_sweep and self.kernel_tail) else 0
        self.L.check(self.lib.kmb_peer_exchange_region(self.ctx, region_index, a, b, self.value + 1, mode, stream), "kmb_peer_exchange_region")

    def join(self):
        stream = torch.cuda.current_stream(self.G.device).cuda_stream
        self.L.check(self.lib.kmb_peer_join(self.ctx, stream), "kmb_peer_join")

    def mark(self):
        self.L.check(self.lib.kmb_peer_mark(self.ctx), "kmb_peer_mark")

    def join_mark(self):
        stream = torch.cuda.current_stream(self.G.device).cuda_stream
        self.L.check(self.lib.kmb_peer_join_mark(self.ctx, stream), "kmb_peer_join_mark")

    def end_step(self):
        self.value += 1

    def __del__(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            try:
                self.lib.kmb_peer_ctx_destroy(ctx)
            except Exception:
                pass


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


def split_pieces(regions, piece):
    """Regions longer than `piece` elements cut into equal pieces (multiples of 64 elements, the last one takes the
    remainder); order, coverage and stages are preserved."""
    out = []
    for a, b, s_ in regions:
        n = max(1, -(-(b - a) // piece))
        step = (-(-(b - a) // n) + 63) // 64 * 64
        out += [(lo, min(b, lo + step), s_) for lo in range(a, b, step)]
    return out


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
        self.peer = None
        self.deferred_stages = set()
        self._marked = False
        if self.world > 1 and st.G.is_cuda and os.environ.get("KMBART_GRAD_EXCHANGE", "peer") != "nccl":
            # pieces of at most PEER_PIECE elements: consecutive pieces run on alternating lanes of the exchange, so the
            # second half of one overlaps the first half of the next (matters for the 160 MB tied-embedding region)
            pieces = split_pieces(self.regions, PEER_PIECE)
            self.peer = PeerExchange.create(st.G, pieces, process_group)
            if self.peer is not None:
                self.regions = pieces
                stages = sorted({s_ for _, _, s_ in pieces if s_ is not None})
                if self.defer_tail:     # the last stage(s) of the sweep cannot finish their exchange before it ends either
                    self.deferred_stages = set(stages[len(stages) - int(os.environ.get("KMBART_DEFER_STAGES", "1")):]) \
                        if int(os.environ.get("KMBART_DEFER_STAGES", "1")) > 0 else set()
                    self.tail_ranges = [(a, b) for a, b, s_ in pieces if s_ is None or s_ in self.deferred_stages]
        self.transport = "peer" if self.peer is not None else ("nccl" if st.G.is_cuda else "gloo")
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
        if os.environ.get("KMBART_GRAD_EXCHANGE") == "none":   # timing experiments: a data-parallel step with no exchange at all
            return 0
        G = self.eng.store.G
        for i, (a, b, s) in enumerate(self.regions):
            if s == stage and self.peer is not None:
                if stage in self.deferred_stages and not self._marked:
                    self.peer.mark()             # finish() joins up to here; the optimizer joins the rest
                    self._marked = True
                self.peer.exchange(i, after_sweep=stage is None or stage in self.deferred_stages)
                self.bytes_reduced += 4 * (b - a)
            elif s == stage:
                self.pending.append(dist.all_reduce(G[a:b], op=dist.ReduceOp.AVG if G.is_cuda else dist.ReduceOp.SUM,
                                                    group=self.group, async_op=True))
                if not G.is_cuda:
                    self._scale_cpu = True
                self.bytes_reduced += 4 * (b - a)
        return 0

    def finish(self):
        if self.peer is not None and self.enabled and self.world > 1:
            if self.defer_tail:
                if not self._marked:
                    self.peer.mark()
                self._marked = False
                self.peer.join_mark()            # everything but the deferred regions
                self.launch_stage(None)
                self.tail_pending = [self.peer]  # joined by AdamW.step() / wait_tail()
            else:
                self.launch_stage(None)
                self.peer.join()
            self.peer.end_step()
            return 0
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
            if w is self.peer:
                w.join()
            else:
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
