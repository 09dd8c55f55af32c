"""AdamW with the semantics of transformers==3.0.2 `AdamW` (what the reference's scripts import:
vcg_train.py:13,100 / pretrain.py:13,100): eps added to the un-corrected sqrt(v), bias correction
folded into the step size, decoupled weight decay applied after the update, defaults
lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True.

The whole update is ONE launch of the multi-tensor kernel (kmb_adamw_multi) over every
parameter; when the parameters belong to a kmbart engine the same launch refreshes the bf16
shadow weights, so the next forward needs no cast pass.  state_dict() keeps the per-parameter
{'step', 'exp_avg', 'exp_avg_sq'} layout of the reference's optimizer checkpoints
(src/utils.py:20-39)."""
import ctypes as C

import torch

from . import lib as L


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self._tables = {}
        self.launches_last = 0

    def _shadow_ptr(self, p):
        """bf16 shadow slot of p when it is a view of an engine's flat buffer, else 0."""
        store = getattr(p, "_kmb_store", None)
        if store is None:
            return 0
        off = p.data_ptr() - store.P.data_ptr()
        if off < 0 or off >= store.total * 4:
            return 0
        return store.P16.data_ptr() + off // 2

    @staticmethod
    def _reducer(p):
        store = getattr(p, "_kmb_store", None)
        return getattr(store, "grad_reducer", None) if store is not None else None

    def _in_tail(self, p):
        red = self._reducer(p)
        if red is None or not getattr(red, "defer_tail", False):
            return False
        store = p._kmb_store
        off = (p.data_ptr() - store.P.data_ptr()) // 4
        return any(a <= off < b for a, b in red.tail_ranges)

    def _ordered(self, group):
        """Parameters with gradients; those whose gradient exchange may still be in flight when step() starts
        (FlatGradReducer(defer_tail=True)) go last: step() launches the head, joins the exchange, then the tail."""
        plist = [p for p in group["params"] if p.grad is not None]
        tail = [self._in_tail(p) for p in plist]
        if not any(tail):
            return plist
        return [p for p, t in zip(plist, tail) if not t] + [p for p, t in zip(plist, tail) if t]

    def _build_table(self, gi, group):
        lib = L.load()
        chunk = lib.kmb_adamw_chunk_elems()
        plist = self._ordered(group)
        rows, cmap = [], []
        n_head_chunks = None
        for ti, p in enumerate(plist):
            if n_head_chunks is None and self._in_tail(p):
                n_head_chunks = len(cmap) // 2
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p.data)
                st["exp_avg_sq"] = torch.zeros_like(p.data)
            if p.grad.is_sparse:
                raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
            assert p.data.is_cuda and p.data.dtype == torch.float32 and p.data.is_contiguous() and p.grad.is_contiguous()
            n = p.numel()
            rows += [p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                     self._shadow_ptr(p), n]
            for c in range((n + chunk - 1) // chunk):
                cmap += [ti, c]
        dev = plist[0].device
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        cm = torch.tensor(cmap, dtype=torch.int32).to(dev)
        step0 = max((self.state[p]["step"] for p in plist), default=0)
        step_dev = torch.tensor([step0, 0], dtype=torch.int32, device=dev)   # {step, float step_size}
        sig = tuple(rows)
        n_chunks = len(cmap) // 2
        return {"sig": sig, "table": table, "cmap": cm, "n_chunks": n_chunks, "step": step_dev, "plist": plist,
                "n_head": n_chunks if n_head_chunks is None else n_head_chunks}

    def _signature(self, group):
        sig = []
        for p in self._ordered(group):
            st = self.state.get(p, {})
            ea, es = st.get("exp_avg"), st.get("exp_avg_sq")
            sig += [p.data_ptr(), p.grad.data_ptr(), ea.data_ptr() if ea is not None else 0,
                    es.data_ptr() if es is not None else 0, self._shadow_ptr(p), p.numel()]
        return tuple(sig)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        L.require_b200()
        self.launches_last = 0
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            t = self._tables.get(gi)
            if t is None or t["sig"] != self._signature(group):
                t = self._build_table(gi, group)
                self._tables[gi] = t
            stream = torch.cuda.current_stream(t["table"].device).cuda_stream
            hyper = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                     float(group["weight_decay"]), int(bool(group["correct_bias"])), 0)
            reducers = {id(r): r for r in (self._reducer(p) for p in t["plist"]) if r is not None}
            n_head, n_all = t["n_head"], t["n_chunks"]
            if n_head in (0, n_all):        # nothing deferred (or everything): one launch
                for r in reducers.values():
                    r.wait_tail()
                L.check(lib.kmb_adamw_multi(t["table"].data_ptr(), t["cmap"].data_ptr(), n_all, t["step"].data_ptr(), *hyper, stream),
                        "kmb_adamw_multi")
                self.launches_last += 2
            else:                           # head while the last all-reduce is in flight, join, tail
                L.check(lib.kmb_adamw_multi_part(t["table"].data_ptr(), t["cmap"].data_ptr(), n_head, t["step"].data_ptr(), *hyper, 1,
                                                 stream), "kmb_adamw_multi_part")
                for r in reducers.values():
                    r.wait_tail()
                L.check(lib.kmb_adamw_multi_part(t["table"].data_ptr(), t["cmap"].data_ptr() + 8 * n_head, n_all - n_head,
                                                 t["step"].data_ptr(), *hyper, 0, stream), "kmb_adamw_multi_part")
                self.launches_last += 3
            for p in t["plist"]:
                self.state[p]["step"] += 1
                store = getattr(p, "_kmb_store", None)
                if store is not None and self._shadow_ptr(p):
                    store.shadow_touched = True
        return loss
