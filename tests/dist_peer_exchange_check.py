"""Run under torchrun (one rank per GPU): the peer-memory gradient exchange (kmbart.parallel.PeerExchange,
csrc/peer_exchange.cu) against NCCL all-reduce on the same buffers — ragged region boundaries, several steps, every
rank must end with bit-identical averages — then times one exchange of a KM-BART-base sized gradient buffer.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_peer_exchange_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
import torch.distributed as dist
from kmbart.parallel import PeerExchange

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
dev = torch.device("cuda", torch.cuda.current_device())

# ---- correctness: ragged regions (unaligned starts, a region shorter than the world size, a large one)
bounds = [0, 3, 1000, 1001 + 7, 300_000, 300_000 + 1_234_567, 5_000_001]
regions = [(bounds[i], bounds[i + 1], i) for i in range(len(bounds) - 1)]
total = bounds[-1]
G = torch.zeros(total, dtype=torch.float32, device=dev)
px = PeerExchange.create(G, regions, None)
assert px is not None, "peer exchange unavailable on this box"
ok = True
for step in range(4):
    g = torch.Generator(device=dev).manual_seed(1000 * step + rank)
    G.copy_(torch.randn(total, device=dev, generator=g) * (1 + rank))
    parts = [torch.empty_like(G) for _ in range(world)]
    dist.all_gather(parts, G)
    want = parts[0].clone()
    for r in range(1, world):
        want += parts[r]                       # rank order, like px_reduce_kernel
    want *= 1.0 / world
    torch.cuda.synchronize()
    order = list(range(len(regions)))
    if step % 2:
        order.reverse()
    for i in order:
        px.exchange(i, after_sweep=(i + step) % 2 == 0)       # both transports (after_sweep selects the one-kernel path when world > 2)
    px.join()
    px.end_step()
    torch.cuda.synchronize()
    same = torch.equal(G, want)
    every = [torch.empty_like(G) for _ in range(world)]
    dist.all_gather(every, G)
    ident = all(torch.equal(every[0], e) for e in every)
    if rank == 0:
        print(f"step {step}: equals rank-ordered mean {same}, identical on every rank {ident}, max err {(G - want).abs().max().item():.3e}")
    ok = ok and same and ident
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
assert int(flag) == 1, "peer exchange mismatch"
del px

# ---- timing: 140 M parameters cut like the model (12 layer regions + a 160 MB tail)
layer = 7_100_000
sizes = [layer] * 12 + [40_000_000]
offs = [0]
for s_ in sizes:
    offs.append(offs[-1] + s_)
regions = [(offs[i], offs[i + 1], i) for i in range(len(sizes))]
G = torch.randn(offs[-1], dtype=torch.float32, device=dev)
px = PeerExchange.create(G, regions, None)
G2 = G.clone()

def peer_all():
    for i in range(len(regions)):
        px.exchange(i, after_sweep=i == len(regions) - 1)
    px.join()
    px.end_step()

def nccl_all():
    ws = [dist.all_reduce(G2[a:b], op=dist.ReduceOp.AVG, async_op=True) for a, b, _ in regions]
    for w in ws:
        w.wait()

for name, fn in (("peer", peer_all), ("nccl", nccl_all)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = 4 * offs[-1] / 1e9
        print(f"{name}: {float(t):.3f} ms per {gb:.2f} GB exchange, world {world}  (algbw {gb / float(t) * 1e3:.0f} GB/s)")
dist.barrier()
if rank == 0:
    print("PEER_EXCHANGE_OK")
dist.destroy_process_group()
