"""CPU suite for the N > 1 path: world_size-2 gloo run of kmbart.parallel.FlatGradReducer on a real ParamStore
(flat gradient buffer of the small product model) — region plan covers every gradient exactly once, stages follow the
backward order, and the exchange averages across ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases as G
from helpers import product_config


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StubEngine:
    def __init__(self, model, cfg):
        from kmbart.engine import ParamStore
        self.store, self.cfg, self.plans, self.grad_reducer = ParamStore(model), cfg, {}, None


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from src.model.model import MultiModalBartForConditionalGeneration
        from kmbart.parallel import FlatGradReducer
        torch.manual_seed(100 + rank)               # different initial weights per rank on purpose
        cfg = product_config(G.small_config())
        model = MultiModalBartForConditionalGeneration(cfg)
        eng = _StubEngine(model, cfg)
        red = FlatGradReducer(model, engine=eng)     # broadcasts rank 0's parameters
        st = eng.store
        w0 = st.P.clone()
        gathered = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        # regions: exact cover, decoder stages before encoder stages, tied embedding last
        cover = torch.zeros(st.total, dtype=torch.int32)
        for a, b, s in red.regions:
            cover[a:b] += 1
        assert bool((cover == 1).all())
        stages = [s for _, _, s in red.regions if s is not None]
        assert sorted(stages) == list(range(cfg.decoder_layers + cfg.encoder_layers))
        assert red.regions[-1][2] is None and red.regions[-1][0] <= st.offsets["model.shared.weight"]
        # exchange: grads = rank + 1 everywhere -> average 1.5; staged launches then finish()
        st.G.fill_(float(rank + 1))
        for s in range(cfg.decoder_layers + cfg.encoder_layers):
            red.launch_stage(s)
        red.finish()
        assert torch.allclose(st.G, torch.full_like(st.G, (1 + world) / 2))
        assert red.bytes_reduced == 4 * st.total
        red.bytes_reduced = 0
        # defer_tail: on a CPU group nothing is left in flight (finish() still joins everything) and wait_tail() is a no-op;
        # the deferred ranges are exactly the stage-None regions (small tensors + tied embedding)
        red.defer_tail = True
        st.G.fill_(float(2 * rank + 1))
        for s in range(cfg.decoder_layers + cfg.encoder_layers):
            red.launch_stage(s)
        red.finish()
        assert red.tail_pending == []
        red.wait_tail()
        assert torch.allclose(st.G, torch.full_like(st.G, float(world)))          # mean of 1, 3 -> 2 at world 2
        assert red.tail_ranges == [(a, b) for a, b, s_ in red.regions if s_ is None] and len(red.tail_ranges) >= 1
        assert any(a <= st.offsets["model.shared.weight"] < b for a, b in red.tail_ranges)
        red.defer_tail = False
        # no_sync(): accumulation micro-steps leave local gradients untouched
        st.G.fill_(float(rank + 1))
        with red.no_sync():
            red.launch_stage(0)
            red.finish()
        assert torch.allclose(st.G, torch.full_like(st.G, float(rank + 1)))
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_region_plan_is_contiguous_and_ordered():
    from kmbart.parallel import plan_regions, layer_stage
    assert layer_stage("model.decoder.layers.5.fc1.weight", 6, 6) == 0
    assert layer_stage("model.decoder.layers.0.fc1.weight", 6, 6) == 5
    assert layer_stage("model.encoder.layers.5.fc1.weight", 6, 6) == 6
    assert layer_stage("model.shared.weight", 6, 6) is None
    names = ["b", "enc0", "dec0", "shared"]
    offs = {"b": 0, "enc0": 64, "dec0": 192, "shared": 320}
    nums = {"b": 10, "enc0": 100, "dec0": 100, "shared": 50}
    import kmbart.parallel as P
    orig = P.layer_stage
    P.layer_stage = lambda n, nd, ne: {"enc0": 1, "dec0": 0}.get(n)
    try:
        regs = plan_regions(names, offs, nums, 10, 384, 1, 1)
    finally:
        P.layer_stage = orig
    assert regs == [(0, 10, None), (10, 164, 1), (164, 292, 0), (292, 384, None)]


def test_exchange_pieces_cover_every_region_once():
    """kmbart.parallel.split_pieces: the pieces of the peer-memory exchange partition each region, keep its stage and
    order, start on 64-element boundaries relative to the region start, and never exceed the piece size by more than the
    64-element rounding."""
    from kmbart.parallel import split_pieces
    regions = [(0, 10, None), (10, 7_100_010, 3), (7_100_010, 7_100_011, 2), (7_100_011, 50_000_000, None)]
    piece = 12 * 1024 * 1024
    out = split_pieces(regions, piece)
    pos, ri = 0, 0
    for a, b, s_ in out:
        assert a == pos and b > a
        while not (regions[ri][0] <= a and b <= regions[ri][1]):
            ri += 1
        assert s_ == regions[ri][2]
        assert (a - regions[ri][0]) % 64 == 0
        assert b - a <= piece + 64
        pos = b
    assert pos == regions[-1][1]
    assert [r for r in out if r[2] == 3] == [(10, 7_100_010, 3)]                  # below the piece size: untouched
    assert len([r for r in out if r[0] >= 7_100_011]) == 4                        # 42.9 M elements -> 4 pieces
    assert split_pieces([(0, 5, 1)], 2) == [(0, 5, 1)]                            # rounding to 64 never produces empty pieces
