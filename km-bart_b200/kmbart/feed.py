"""Batch feed: collator output -> device.

replaces: the per-tensor `.to(device)` loop of the reference's training / generation loops
  (src/training.py:120-130: `image_features=list(map(lambda x: x.to(device), batch['image_features']))` — B separate
  synchronous host->device copies of [n_i, 2052] fp32 tensors, 295 KB per sample — plus one `.to(device)` per id /
  mask tensor; src/generation.py:22-32 does the same).

`DeviceFeeder` copies a collated host batch on a side stream while the previous step computes:
  * every RoI tensor of the list is copied straight into its row range of ONE device staging buffer
    `[R_total, 2052]` (no host-side packing pass), the id / mask / label tensors into theirs;
  * two staging slots rotate, guarded by CUDA events in both directions (the compute stream waits for the copy, the
    next copy into a slot waits for the step that read it), so no host synchronisation is involved;
  * the model receives the features as `PackedImageFeatures` (device tensor + per-sample counts), which the engine
    consumes without the pointer-table gather (`Engine._stage_inputs` packed path).
The feeder needs the sm_100 engine's consumer but launches no kernels of its own; there is no CPU fallback: it raises
if CUDA is unavailable.
"""
import torch


class PackedImageFeatures:
    """RoI features of a batch as one device tensor [sum(n_i), 2052] plus the per-sample row counts.  Behaves like the
    reference's list of per-sample tensors for code that only iterates / indexes (views, no copies)."""

    def __init__(self, tensor, counts):
        self.tensor = tensor
        self.counts = [int(c) for c in counts]
        assert tensor.dim() == 2 and tensor.shape[0] == sum(self.counts)

    def __len__(self):
        return len(self.counts)

    def __getitem__(self, i):
        off = sum(self.counts[:i])
        return self.tensor[off:off + self.counts[i]]

    def __iter__(self):
        off = 0
        for c in self.counts:
            yield self.tensor[off:off + c]
            off += c

    def as_list(self):
        return list(self)


def unwrap_features(image_features, image_counts=None):
    """(features, counts) for the engine: PackedImageFeatures -> (tensor, counts); anything else unchanged."""
    if isinstance(image_features, PackedImageFeatures):
        return image_features.tensor, image_features.counts
    return image_features, image_counts


class _Slot:
    def __init__(self):
        self.bufs = {}           # key -> device tensor
        self.ready = None        # copy finished (recorded on the copy stream)
        self.released = None     # consumer finished (recorded on the compute stream)


class DeviceFeeder:
    """put(host_batch) starts the asynchronous copy of a collated batch; get() returns the device batch of the oldest
    put(), ordered after its copy on the current stream.  Host tensors should be pinned (`pin_memory()`), otherwise
    the copies are synchronous with respect to the host, as in torch."""

    FEATURE_KEY = "image_features"

    def __init__(self, device, depth=2):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceFeeder needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(max(2, int(depth)))]
        self.n_put = 0
        self.n_get = 0
        self.pending = []        # (slot index, device batch)
        self.bytes_last = 0

    def _buf(self, slot, key, shape, dtype):
        t = slot.bufs.get(key)
        if t is None or t.dtype != dtype or t.numel() < int(torch.Size(shape).numel()):
            t = torch.empty(int(torch.Size(shape).numel()), dtype=dtype, device=self.device)
            slot.bufs[key] = t
        return t[:int(torch.Size(shape).numel())].view(shape)

    def put(self, host_batch):
        if len(self.pending) >= len(self.slots):
            raise RuntimeError("DeviceFeeder: every staging slot is in flight; call get() first")
        idx = self.n_put % len(self.slots)
        slot = self.slots[idx]
        self.n_put += 1
        out, nbytes = {}, 0
        with torch.cuda.stream(self.stream):
            if slot.released is not None:
                self.stream.wait_event(slot.released)   # the step that read this slot has been enqueued and must finish first
            for key, val in host_batch.items():
                if isinstance(val, (list, tuple)) and key == self.FEATURE_KEY:
                    counts = [int(f.shape[0]) for f in val]
                    width = int(val[0].shape[1]) if val else 0
                    dst = self._buf(slot, key, (sum(counts), width), val[0].dtype if val else torch.float32)
                    off = 0
                    for f, c in zip(val, counts):
                        if c:
                            dst[off:off + c].copy_(f, non_blocking=True)
                            nbytes += f.numel() * f.element_size()
                        off += c
                    out[key] = PackedImageFeatures(dst, counts)
                elif isinstance(val, torch.Tensor):
                    dst = self._buf(slot, key, tuple(val.shape), val.dtype)
                    dst.copy_(val, non_blocking=True)
                    nbytes += val.numel() * val.element_size()
                    out[key] = dst
                else:   # host-side objects (e.g. relation label dicts) pass through untouched
                    out[key] = val
            slot.ready = torch.cuda.Event()
            slot.ready.record(self.stream)
        self.bytes_last = nbytes
        self.pending.append((idx, out))
        return nbytes

    def get(self):
        if not self.pending:
            raise RuntimeError("DeviceFeeder: get() without a pending put()")
        idx, out = self.pending.pop(0)
        slot = self.slots[idx]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot.ready)
        self.n_get += 1
        self._last_slot = slot
        return out

    def release(self):
        """Marks the batch returned by the last get() as consumed: call after the step that uses it has been enqueued
        (its kernels read the staging buffers in stream order)."""
        slot = getattr(self, "_last_slot", None)
        if slot is not None:
            slot.released = torch.cuda.Event()
            slot.released.record(torch.cuda.current_stream(self.device))

    def __call__(self, batches):
        """Generator over device batches with one batch of look-ahead: `for b in feeder(loader): step(b)`."""
        it = iter(batches)
        try:
            self.put(next(it))
        except StopIteration:
            return
        while self.pending:
            b = self.get()
            try:
                self.put(next(it))
            except StopIteration:
                pass
            yield b
            self.release()
