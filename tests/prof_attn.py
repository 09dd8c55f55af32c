"""Standalone driver for the attention kernels at the config-2 shapes (for ncu and CUDA-event timing).
   python tests/prof_attn.py [enc|dec|cross] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from kmbart import lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "enc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, H = 128, 12
Sq, Sk, causal = {"enc": (100, 100, 0), "dec": (48, 48, 1), "cross": (48, 100, 0)}[which]
d = H * 64
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device="cuda", generator=g).to(torch.bfloat16)
q, k, v, do = rnd(B * Sq, d), rnd(B * Sk, d), rnd(B * Sk, d), rnd(B * Sq, d)
o = torch.zeros(B * Sq, d, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B * H * Sq, device="cuda")
dscr = torch.zeros(B * H * Sq, device="cuda")
dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def fwd():
    L.check(lib.kmb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), d, d, d, o.data_ptr(), d, lse.data_ptr(), 0, B, H, Sq, Sk, 64, causal, 0.125, st), "fwd")
def bwd():
    L.check(lib.kmb_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), d, d, d, o.data_ptr(), d, do.data_ptr(), d, lse.data_ptr(), dscr.data_ptr(), 0,
                             dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), d, d, d, B, H, Sq, Sk, 64, causal, 0.125, st), "bwd")
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    io = (4 if name == "fwd" else 8) * B * (Sq + Sk) / 2 * d * 2
    print(f"{which} {name}: {1e3 * tot / reps:.1f} us  ({io / (tot / reps * 1e-3) / 1e9:.0f} GB/s algorithmic)")
