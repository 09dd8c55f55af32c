"""Debug driver (not a pytest): small-config fused train step vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "km-bart_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import kmbart_oracle as O
from helpers import small_config, product_config, load_oracle_weights, to_cuda_batch, rel_err
from src.model.model import MultiModalBartForConditionalGeneration

torch.manual_seed(0)
ragged = len(sys.argv) > 1 and sys.argv[1] == "ragged"
ocfg = small_config()
sd = O.init_state_dict(ocfg, seed=0)
# make LN params / biases non-trivial so their gradients are exercised
g = torch.Generator().manual_seed(5)
for k in sd:
    if k.endswith(".bias") or "layer_norm" in k or "layernorm" in k:
        sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
sd["final_logits_bias"] = 0.1 * torch.randn(1, ocfg.vocab_size, generator=g)
batch = O.synthetic_batch(ocfg, batch=4, n_regions=6, n_ctx=14, tgt_len=10, seed=3, ragged=ragged)

# oracle with autograd
osd = {k: v.clone().requires_grad_(k != "final_logits_bias") for k, v in sd.items()}
loss_o, logits_o, h_o, enc_o = O.forward_conditional_generation(osd, ocfg, **batch)
loss_o.backward()
print("oracle loss", loss_o.item())

model = MultiModalBartForConditionalGeneration(product_config(ocfg))
load_oracle_weights(model, sd)
model.cuda().train()
cb = to_cuda_batch(batch)
out = model(**cb)
loss = out[0]
print("kernel loss", loss.item(), "rel", abs(loss.item() - loss_o.item()) / abs(loss_o.item()))
print("enc rel err", rel_err(out[2], enc_o))
a = model._engine().last_train[0]
print("dec rel err", rel_err(a["dec_f32"].view(h_o.shape), h_o))
logits = out[1].materialize()
print("logits max abs err", (logits.float().cpu() - logits_o).abs().max().item(), "rel", rel_err(logits, logits_o))
loss.backward()
torch.cuda.synchronize()
worst = []
for n, p in model.named_parameters():
    go = osd[n].grad
    gk = p.grad
    if gk is None:
        print("NO GRAD", n); continue
    e = rel_err(gk, go)
    worst.append((e, n, go.norm().item()))
worst.sort(reverse=True)
for e, n, nrm in worst[:25]:
    print(f"grad rel err {e:.4f}  |g|={nrm:.3e}  {n}")
print("median grad rel err", sorted(w[0] for w in worst)[len(worst) // 2])
