"""Micro-timer of the bandwidth-bound kernels at the config-2 shapes (CUDA events, rotating buffers larger than L2).
   python tests/time_kernels.py [ln_bwd|ln_fwd|attn|colsum ...]   prints us per launch and achieved GB/s vs MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from kmbart import lib as L

L.require_b200()
lib = L.load()
BF16, F32 = torch.bfloat16, torch.float32
PEAK = 6500.0
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
which = set(sys.argv[1:]) or {"ln_bwd", "ln_fwd", "attn", "colsum"}
st = torch.cuda.current_stream().cuda_stream
seed = torch.tensor([1234], dtype=torch.int64, device="cuda")


def timeit(fn, nbuf, iters=20):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def report(name, us, nbytes):
    print(f"{name:44s} {us:8.1f} us   {nbytes / us / 1e3:7.0f} GB/s   {nbytes / us / 1e3 / PEAK:5.2f} of HBM peak", flush=True)


d = 768
for M, tag in ((12800, "enc"), (6144, "dec")):
    nbuf = 3
    if "ln_bwd" in which:
        dyA = [torch.randn(M, d, device="cuda") for _ in range(nbuf)]
        dyB = [torch.randn(M, d, device="cuda").to(BF16) for _ in range(nbuf)]
        pre = [torch.randn(M, d, device="cuda") for _ in range(nbuf)]
        dpre = [torch.empty(M, d, device="cuda") for _ in range(nbuf)]
        dz = [torch.empty(M, d, device="cuda", dtype=BF16) for _ in range(nbuf)]
        mean, rstd = torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")
        gamma = torch.ones(d, device="cuda")
        dg, db, dbias = (torch.zeros(d, device="cuda") for _ in range(3))
        for pdrop in (0.0, 0.1):
            def f(i):
                L.check(lib.kmb_layernorm_bwd(dyA[i].data_ptr(), dyB[i].data_ptr(), pre[i].data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              gamma.data_ptr(), dpre[i].data_ptr(), dz[i].data_ptr(), dg.data_ptr(), db.data_ptr(),
                                              dbias.data_ptr(), M, d, 0.0, 0, pdrop, 7, seed.data_ptr(), st))
            report(f"ln_bwd {tag} M={M} drop_out={pdrop}", timeit(f, nbuf), M * d * 16)
        del dyA, dyB, pre, dpre, dz
    if "ln_fwd" in which:
        z = [torch.randn(M, d, device="cuda").to(BF16) for _ in range(nbuf)]
        res = [torch.randn(M, d, device="cuda") for _ in range(nbuf)]
        pre = [torch.empty(M, d, device="cuda") for _ in range(nbuf)]
        y32 = [torch.empty(M, d, device="cuda") for _ in range(nbuf)]
        y16 = [torch.empty(M, d, device="cuda", dtype=BF16) for _ in range(nbuf)]
        mean, rstd = torch.zeros(M, device="cuda"), torch.ones(M, device="cuda")
        gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
        def f(i):
            L.check(lib.kmb_layernorm_fwd(z[i].data_ptr(), res[i].data_ptr(), gamma.data_ptr(), beta.data_ptr(), pre[i].data_ptr(),
                                          y32[i].data_ptr(), y16[i].data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, d, 0.1, 5,
                                          seed.data_ptr(), st))
        report(f"ln_fwd {tag} M={M} drop=0.1", timeit(f, nbuf), M * d * 16)
        del z, res, pre, y32, y16
    if "colsum" in which:
        x = [torch.randn(M, 3 * d, device="cuda").to(BF16) for _ in range(nbuf)]
        out = torch.zeros(3 * d, device="cuda")
        def f(i):
            L.check(lib.kmb_colsum_bf16(x[i].data_ptr(), 3 * d, out.data_ptr(), M, 3 * d, st))
        report(f"colsum {tag} [{M},{3 * d}]", timeit(f, nbuf), M * 3 * d * 2)
        del x

if "attn" in which:
    B, H = 128, 12
    for Sq, Sk, causal, tag in ((100, 100, 0, "enc self"), (48, 48, 1, "dec self"), (48, 100, 0, "cross")):
        nbuf = 4
        Mq, Mk = B * Sq, B * Sk
        if Sq == Sk:
            qkv = [torch.randn(Mq, 3 * d, device="cuda").to(BF16) for _ in range(nbuf)]
            q = qkv; k = [t[:, d:] for t in qkv]; v = [t[:, 2 * d:] for t in qkv]
            ldq = ldk = ldv = 3 * d
            dqkv = [torch.empty(Mq, 3 * d, device="cuda", dtype=BF16) for _ in range(nbuf)]
            dq = dqkv; dk = [t[:, d:] for t in dqkv]; dv = [t[:, 2 * d:] for t in dqkv]
        else:
            q = [torch.randn(Mq, d, device="cuda").to(BF16) for _ in range(nbuf)]
            kv = [torch.randn(Mk, 2 * d, device="cuda").to(BF16) for _ in range(nbuf)]
            k = kv; v = [t[:, d:] for t in kv]
            ldq, ldk, ldv = d, 2 * d, 2 * d
            dq = [torch.empty(Mq, d, device="cuda", dtype=BF16) for _ in range(nbuf)]
            dkv = [torch.empty(Mk, 2 * d, device="cuda", dtype=BF16) for _ in range(nbuf)]
            dk = dkv; dv = [t[:, d:] for t in dkv]
        o = [torch.empty(Mq, d, device="cuda", dtype=BF16) for _ in range(nbuf)]
        do = [torch.randn(Mq, d, device="cuda").to(BF16) for _ in range(nbuf)]
        lse = torch.empty(B * H * Sq, device="cuda")
        dscr = torch.empty(B * H * Sq, device="cuda")
        def ffwd(i):
            L.check(lib.kmb_attn_fwd(q[i].data_ptr(), k[i].data_ptr(), v[i].data_ptr(), ldq, ldk, ldv, o[i].data_ptr(), d, lse.data_ptr(),
                                     0, B, H, Sq, Sk, 64, causal, 0.125, st))
        fwd_bytes = (2 * Mq + 2 * Mk) * d * 2
        report(f"attn_fwd {tag} Sq={Sq} Sk={Sk}", timeit(ffwd, nbuf), fwd_bytes)
        def fbwd(i):
            L.check(lib.kmb_attn_bwd(q[i].data_ptr(), k[i].data_ptr(), v[i].data_ptr(), ldq, ldk, ldv, o[i].data_ptr(), d, do[i].data_ptr(), d,
                                     lse.data_ptr(), dscr.data_ptr(), 0, dq[i].data_ptr(), dk[i].data_ptr(), dv[i].data_ptr(),
                                     ldq, ldk, ldv, B, H, Sq, Sk, 64, causal, 0.125, st))
        bwd_bytes = (Mq + 2 * Mk) * d * 2 * 2 + 2 * Mq * d * 2
        report(f"attn_bwd {tag} Sq={Sq} Sk={Sk}", timeit(fbwd, nbuf), bwd_bytes)
