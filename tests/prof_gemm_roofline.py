"""The GEMM launch bench.py reports under `roofline` (fc1 shape 12800x3072x768, bf16 out), for ncu:
   ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 3 -c 1 -o X python tests/prof_gemm_roofline.py [gelu]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from kmbart import lib as L
L.require_b200()
lib = L.load()
M, N, K = 12800, 3072, 768
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e = L.GemmEpilogue()
e.alpha, e.out_bf16, e.ld_bf16 = 1.0, out.data_ptr(), N
if len(sys.argv) > 1 and sys.argv[1] == "gelu":
    e.act, e.out_preact, e.bias = L.ACT_GELU, pre.data_ptr(), bias.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for i in range(6):
    flush.zero_()
    L.check(lib.kmb_gemm(A.data_ptr(), W.data_ptr(), M, N, K, K, K, 0, 0, 0, C.byref(e), 0, st), "gemm")
torch.cuda.synchronize()
print("done")
