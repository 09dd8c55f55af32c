"""2-GPU numerical check of the data-parallel path (run under torchrun, 2 ranks, NCCL):
  * FlatGradReducer (all-reduce points inside the backward plan, deferred tail) and torch DDP (what the reference's
    scripts use, vcg_train.py:96-98) must both give every rank the gradient of the CONCATENATED batch computed on one GPU
    (mean over the same number of target tokens per rank), and the same weights after an AdamW step;
  * losses: mean of the rank losses == single-GPU loss of the concatenated batch.
Exit code 0 = all checks passed (tests/test_gpu_multi.py asserts it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "km-bart_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from oracle import kmbart_oracle as O
import golden_cases as G
from helpers import product_config, load_oracle_weights, to_cuda_batch, rel_err
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.optim import AdamW
from kmbart.parallel import FlatGradReducer


def make(ocfg, sd):
    m = MultiModalBartForConditionalGeneration(product_config(ocfg))
    load_oracle_weights(m, sd)
    return m.cuda().train()


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    ocfg = G.small_config()
    sd = G.perturb(O.init_state_dict(ocfg, seed=0))
    full = O.synthetic_batch(ocfg, batch=4 * world, n_regions=6, n_ctx=14, tgt_len=10, seed=3)   # equal token counts per row
    sl = slice(4 * rank, 4 * rank + 4)
    mine = {k: (v[sl] if not isinstance(v, list) else v[sl]) for k, v in full.items()}
    # single-GPU reference on the concatenated batch (every rank computes it, identical)
    ref = make(ocfg, sd)
    opt_ref = AdamW(ref.parameters(), lr=1e-3)
    l_ref = ref(**to_cuda_batch(full))[0]
    l_ref.backward()
    g_ref = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
    opt_ref.step()
    w_ref = {n: p.detach().clone() for n, p in ref.named_parameters()}
    worst = {}
    for mode in ("flat", "ddp"):
        m = make(ocfg, sd)
        step_model = m
        if mode == "flat":
            m._engine()
            FlatGradReducer(m, defer_tail=True)
        else:
            m._engine()
            from torch.nn.parallel import DistributedDataParallel as DDP
            step_model = DDP(m, device_ids=[torch.cuda.current_device()], find_unused_parameters=True)
        opt = AdamW(m.parameters(), lr=1e-3)
        loss = step_model(**to_cuda_batch(mine))[0]
        opt.zero_grad()
        loss.backward()
        if mode == "flat":
            m._engine().grad_reducer.wait_tail()
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        assert abs(lt.item() / world - l_ref.item()) <= 1e-3 * abs(l_ref.item()), (mode, lt.item() / world, l_ref.item())
        w = 0.0
        for n, p in m.named_parameters():
            if g_ref[n].norm() < 1e-7:
                continue
            w = max(w, rel_err(p.grad, g_ref[n]))
        assert w <= 2e-2, (mode, "grad", w)
        opt.step()
        torch.cuda.synchronize()
        ww = max(rel_err(p.detach(), w_ref[n]) for n, p in m.named_parameters())
        assert ww <= 1e-3, (mode, "weights", ww)
        # every rank holds the same weights
        for n, p in m.named_parameters():
            t = p.detach().clone()
            dist.broadcast(t, 0)
            assert torch.equal(t, p.detach()), (mode, n)
        worst[mode] = (w, ww)
        del m, opt, step_model
    if rank == 0:
        print("dist_numeric_check ok:", {k: (round(a, 5), round(b, 7)) for k, (a, b) in worst.items()}, "world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
