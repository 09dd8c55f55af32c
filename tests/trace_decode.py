"""In-kernel timeline of the persistent decode step: every CTA stamps %globaltimer when it arrives at and when it is
released from each grid barrier (KmbDecodeStep.trace).  Prints, per barrier-delimited phase, the work time (release of the
previous barrier -> arrival, median / max over CTAs), the barrier latency (last arrival -> first release) and the release
spread, for rows 64 (greedy) and rows 320 (beam 5) at the configs[3] per-GPU shape.
    python tests/trace_decode.py [out.json]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))
import torch
from src.model.config import MultiModalBartConfig
from src.model.model import MultiModalBartForConditionalGeneration
from kmbart.decode import get_session
from kmbart import lib as L

cfg = MultiModalBartConfig.from_dict(json.load(open(os.path.join(ROOT, "configs", "vcg_base.json"))))
torch.manual_seed(0)
model = MultiModalBartForConditionalGeneration(cfg).cuda().eval()
eng = model._engine()
eng.sync_shadow()
lib = L.load()
CL = os.environ.get("KMBART_DECODE_CLUSTER", "0") == "1"
grid = lib.kmb_decode_cluster_grid(cfg.d_model) if CL else lib.kmb_decode_step_grid()
report = {}
for label, rows, use_tbl in (("rows64", 64, False), ("rows320", 320, True)):
    sess = get_session(eng, 64, 100, rows, 25, False)
    sess.begin(torch.randn(64, 100, cfg.d_model, device="cuda") * 0.5, None, cfg.decoder_start_token_id, use_tbl)
    sess.flb = model.final_logits_bias.reshape(-1)
    nbar = lib.kmb_decode_cluster_barriers(cfg.decoder_layers) if CL else 11 * cfg.decoder_layers
    trace = torch.zeros(nbar * grid * 2 + grid * 96, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for t in list(range(12)) * 3:          # warm: fills the caches up to t = 11 and pages everything in
        args = sess._mega_args(t)
        L.check((lib.kmb_decode_step_cluster if CL else lib.kmb_decode_step)(ctypes.byref(args), stream), "decode_step")
    torch.cuda.synchronize()
    args = sess._mega_args(12)
    args.trace = trace.data_ptr()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check((lib.kmb_decode_step_cluster if CL else lib.kmb_decode_step)(ctypes.byref(args), stream), "decode_step")
    e1.record()
    torch.cuda.synchronize()
    ev = trace[nbar * grid * 2:].view(grid, 6, 16).cpu()
    tr = trace[:nbar * grid * 2].view(nbar, grid, 2).cpu().double()
    used = int((tr[:, 0, 0] > 0).sum())
    tr = tr[:used]
    arrive, release = tr[:, :, 0], tr[:, :, 1]
    t0 = arrive[0].min()
    phases = []
    prev_rel = None
    for i in range(used):
        a, r = arrive[i], release[i]
        start = prev_rel if prev_rel is not None else torch.full_like(a, float(t0))
        work = (a - start)
        phases.append({"i": i, "work_med_us": round(work.median().item() / 1e3, 2), "work_max_us": round(work.max().item() / 1e3, 2),
                       "barrier_us": round((r.min() - a.max()).item() / 1e3, 2), "release_spread_us": round((r.max() - r.min()).item() / 1e3, 2),
                       "phase_us": round(((r.max() - (start.max() if prev_rel is not None else t0))).item() / 1e3, 2)})
        prev_rel = r
    total = (release[-1].max() - t0).item() / 1e3
    report[label] = {"kernel_us_events": round(e0.elapsed_time(e1) * 1e3, 1), "barriers": used, "traced_span_us": round(total, 1),
                     "sum_work_max_us": round(sum(p["work_max_us"] for p in phases), 1),
                     "sum_barrier_us": round(sum(p["barrier_us"] for p in phases), 1),
                     "sum_spread_us": round(sum(p["release_spread_us"] for p in phases), 1), "phases": phases}
    print(label, {k: v for k, v in report[label].items() if k != "phases"})
    for p in phases[:14]:
        print("   ", p)
    # intra-phase events (clock64 cycles of thread 0, layer 1): deltas between consecutive stamps
    names = ["entry", "loads+stats", "staged", "w_ready", "mma", "partials", "cl_sync", "-", "reduced", "post", "pre_bar", "post_bar",
             "att_scores", "att_pv", "-", "-"]
    for blk in (0, 1, 5, 48, 100, 127):
        if blk >= grid:
            continue
        for k, pn in enumerate("ABCDEF"):
            e = ev[blk, k].tolist()
            if e[0] == 0:
                continue
            probes = (e[14], e[15])
            seq = sorted([(names[i], e[i]) for i in range(14) if e[i] > 0], key=lambda x: x[1])
            txt = " ".join(f"{n}+{b - a}" for (_, a), (n, b) in zip(seq[:-1], seq[1:]))
            print(f"    blk {blk:3d} phase {pn}: total {seq[-1][1] - seq[0][1]} cyc | {txt} | L2 probe entry/mid {probes}")
if len(sys.argv) > 1:
    json.dump(report, open(sys.argv[1], "w"), indent=1)
