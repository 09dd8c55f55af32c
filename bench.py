#!/usr/bin/env python
"""bench.py — KM-BART-base VCG fine-tuning step (fwd + bwd + AdamW) throughput, BASELINE.json metric.

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle of the
                                                           # reference's path on the host cores

Workload (BASELINE.json configs[1], SURVEY.md §8d "config 2"): bart-base 6+6, d=768, batch 128 per GPU,
36 RoI x 2052 features, S_e = 100, S_d = 48, bf16 tensor-core compute / fp32 accumulate, dropout 0.1 on,
synthetic data, random-init weights.  One "step" = forward + loss + backward + AdamW update.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "km-bart_b200"))

import torch  # noqa: E402

FLOP_PER_SAMPLE = 56.412e9   # 18.804 GFLOP fwd x 3
B_PER_GPU, R, N_CTX, S_D = 128, 36, 64, 48
METRIC = "train samples/s KM-BART-base VCG fine-tune step (fwd+bwd+AdamW)"
# The other training configurations of BASELINE.json: measured by the same run as short extra legs
# (`workloads` in the JSON line) so that the driver's 1/2/4/8-GPU runs carry them too, or alone with --workload.
WORKLOADS = {
    "vcg": dict(config="configs/vcg_base.json", cls="MultiModalBartForConditionalGeneration", batch=128, R=36, n_ctx=64, tgt=48,
                flop=56.412e9, label="configs[1]: KM-BART base VCG fine-tuning step"),
    "pretrain": dict(config="configs/pretrain_base.json", cls="MultiModalBartForPreTraining", batch=128, R=36, n_ctx=64, tgt=48,
                     flop=77.012e9, label="configs[2]: KM-BART base multitask pre-training step (MLM x5 + MRM + attribute + relation heads, S_d = 86)"),
    "large": dict(config=None, cls="MultiModalBartForConditionalGeneration", batch=64, R=100, n_ctx=256, tgt=48,
                  flop=464.662e9, label="configs[4]: KM-BART large (12+12, d = 1024) VCG step, 100 RoIs + 256 ctx tokens (S_e = 356)"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def make_batch(cfg, seed, device=None, pin=False, batch=B_PER_GPU):
    from kmbart.synth import synthetic_batch   # product-side generator; the oracle is not imported by this arm
    b = synthetic_batch(cfg, batch=batch, n_regions=R, n_ctx=N_CTX, tgt_len=S_D, seed=seed)
    if pin:
        b = {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()) for k, v in b.items()}
    if device is not None:
        b = {k: ([t.to(device) for t in v] if isinstance(v, list) else v.to(device)) for k, v in b.items()}
    return b



GEN_B, GEN_NEW = 64, 24    # BASELINE.json configs[3]: batch 512 over 8 GPUs = 64 samples per GPU, 24 new tokens


def bench_generation(model, cfg, dev, rank, world, dist, hbm_peak, peak_src):
    """KV-cached generation (src/model/mixins.py:33-384 path): tokens/s for greedy, top-k sampling and beam-5 through
    model.generate(), plus the decode-step kernel chain timed alone (CUDA-graph replays) against the HBM roofline."""
    from kmbart.decode import get_session
    model.eval()
    dev_b = make_batch(cfg, 4321 + rank, device=dev, batch=GEN_B)
    host_b = make_batch(cfg, 4321 + rank, pin=True, batch=GEN_B)
    gi = dict(input_ids=dev_b["input_ids"], image_features=dev_b["image_features"], attention_mask=dev_b["attention_mask"])
    L_ = GEN_NEW + 1
    modes = {"greedy": dict(max_length=L_, min_length=L_),
             "sample_top50": dict(max_length=L_, min_length=L_, do_sample=True, top_k=50),
             "beam5": dict(max_length=L_, min_length=L_ - 1, num_beams=5, early_stopping=True)}
    out = {"config": f"configs[3] per-GPU share: batch {GEN_B}, S_e=100, {GEN_NEW} new tokens, KV cache, EOS suppressed until the last step",
           "unit": "tokens/s", "tokens_per_s": {}, "ms_per_call": {}}

    def timed_calls(fn, n):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    with torch.no_grad():
        for name, kw in modes.items():
            toks = model.generate(**gi, **kw)      # warm-up: captures the per-step graphs
            assert toks.shape[0] == GEN_B and toks.shape[1] in (L_ - 1, L_), (name, tuple(toks.shape))   # equal-length beam hypotheses come back without the final EOS
            ms = timed_calls(lambda: model.generate(**gi, **kw), 3)
            out["tokens_per_s"][name] = round(GEN_B * world * GEN_NEW / (ms * 1e-3), 1)
            out["ms_per_call"][name] = round(ms, 3)

        def e2e_call():
            b = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev, non_blocking=True))
                 for k, v in host_b.items() if k in ("input_ids", "image_features", "attention_mask")}
            return model.generate(**b, **modes["greedy"]).cpu()
        e2e_call()
        ms = timed_calls(e2e_call, 3)
        h2d = sum((sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size())
                  for k, v in host_b.items() if k in ("input_ids", "image_features", "attention_mask"))
        out["e2e"] = {"value": round(GEN_B * world * GEN_NEW / (ms * 1e-3), 1), "unit": "tokens/s", "mode": "greedy",
                      "h2d_bytes_per_call": int(h2d), "d2h_bytes_per_call": GEN_B * L_ * 8, "ms_per_call": round(ms, 3)}

        # decode-step chain alone: 24 graph replays (greedy: rows 64 incl. device-side selection; beam: rows 320, model chain only)
        eng = model._engine()
        step = {}
        for label, rows, sel, use_tbl in (("rows64_greedy", GEN_B, dict(do_sample=False, temperature=1.0, top_k=50, eos=cfg.eos_token_id,
                                                                     pad=cfg.pad_token_id, min_length=L_), False),
                                          ("rows320_beam5", GEN_B * 5, None, True)):
            sess = get_session(eng, GEN_B, R + N_CTX, rows, L_, False)
            enc = torch.zeros(GEN_B, R + N_CTX, cfg.d_model, device=dev)
            sess.begin(enc, None, cfg.decoder_start_token_id, use_tbl)

            def run_steps():
                for t in range(GEN_NEW):
                    sess.step(t, model.final_logits_bias, sel)
            run_steps()
            ms = timed_calls(run_steps, 5)
            us = ms * 1e3 / GEN_NEW
            nbytes = sum(sess.step_bytes(t) for t in range(GEN_NEW)) / GEN_NEW
            step[label] = {"us_per_step": round(us, 1), "algorithmic_bytes_per_step": int(nbytes),
                           "achieved_GBps": round(nbytes / (us * 1e-6) / 1e9, 1), "frac_of_hbm_peak": round(nbytes / (us * 1e-6) / 1e9 / hbm_peak, 4),
                           "kernels_per_step": sess.launches_per_step}
        out["decode_step"] = step
        out["hbm_peak_GBps"] = hbm_peak
        out["peak_source"] = peak_src
    model.train()
    return out


def _exchange_text(reducer):
    if reducer.transport == "peer":
        return ("FlatGradReducer over NVLink peer memory (csrc/peer_exchange.cu): per-layer regions averaged during the backward sweep by "
                "copy-engine pushes + one small reduction kernel, the last regions (tied embedding, small tensors, encoder layer 0) "
                + ("by one load/store kernel" if reducer.peer.kernel_tail else "the same way") + " while AdamW updates the others")
    return ("FlatGradReducer: per-layer NCCL all-reduce (AVG) overlapped with backward, last region overlapped with the AdamW update of the others")


def bench_workload(name, dev, rank, world, dist, steps, sustained, e2e=False):
    """One of the other BASELINE configs, device-resident inputs, fwd + bwd + AdamW, CUDA events, max over ranks."""
    import gc
    from src.model.config import MultiModalBartConfig
    import src.model.model as M
    from kmbart.optim import AdamW
    from kmbart.synth import synthetic_batch, synthetic_pretrain_batch, to_device
    w = WORKLOADS[name]
    if w["config"]:
        with open(os.path.join(ROOT, w["config"])) as f:
            cfg = MultiModalBartConfig.from_dict(json.load(f))
    else:
        cfg = MultiModalBartConfig()      # src/model/config.py:12-18 defaults = bart-large
    torch.manual_seed(0)
    model = getattr(M, w["cls"])(cfg).to(dev).train()
    if world > 1:
        from kmbart.parallel import FlatGradReducer
        model._engine()
        FlatGradReducer(model, defer_tail=True)
    opt = AdamW(model.parameters(), lr=1e-5)
    gen = synthetic_pretrain_batch if name == "pretrain" else synthetic_batch
    batch = to_device(gen(cfg, batch=w["batch"], n_regions=w["R"], n_ctx=w["n_ctx"], tgt_len=w["tgt"], seed=1234 + rank), dev)

    def step():
        out = model(**batch)
        loss = out[0]["loss"] if isinstance(out[0], dict) else out[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = w["batch"] * world * steps / (ms * 1e-3)
    res = {"workload": w["label"], "value": round(value, 1), "unit": "samples/s", "ms_per_step": round(ms / steps, 3), "steps": steps,
           "batch_per_gpu": w["batch"], "n_gpus": world, "step_mfu": round(value / world * w["flop"] / 1e12 / sustained, 4),
           "loss": float(loss.item()), "gpu_launches": int(model._engine().launches_last + opt.launches_last)}
    if e2e:
        host = to_device(gen(cfg, batch=w["batch"], n_regions=w["R"], n_ctx=w["n_ctx"], tgt_len=w["tgt"], seed=1234 + rank), None, pin=True)
        h2d = sum((sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) and v and torch.is_tensor(v[0]) else
                   (v.numel() * v.element_size() if torch.is_tensor(v) else 0)) for v in host.values())

        def e2e_step():
            b = {k: (v if k == "relation_labels" else ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev, non_blocking=True)))
                 for k, v in host.items()}
            out = model(**b)
            l_ = out[0]["loss"] if isinstance(out[0], dict) else out[0]
            opt.zero_grad()
            l_.backward()
            opt.step()
            return l_.item()
        e2e_step()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            e2e_step()
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = t.item()
        res["e2e"] = {"value": round(w["batch"] * world * steps / (ms2 * 1e-3), 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": 4, "ms_per_step": round(ms2 / steps, 3),
                      "api": "pinned host batch -> non_blocking .to(device) -> model.forward(**batch) + loss.backward() + AdamW.step(), loss.item() each step"}
    del model, opt, batch
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_other_workload(args):
    """--workload pretrain | large as the headline of the JSON line (same contract; e2e = pinned host batch copied every
    step with non_blocking .to(device) on the compute stream + loss.item())."""
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    burst, sustained, hbm, peak_src = load_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = bench_workload(args.workload, dev, rank, world, dist, args.steps, sustained, e2e=True)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
        w = WORKLOADS[args.workload]
        line = {"metric": "train samples/s " + w["label"], "value": res["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": 3, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": w["label"], "global_batch": w["batch"] * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set far exceeds the 126 MB L2"},
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"] * args.steps,
                "roofline": {"bound": "tensor", "kernel": "whole step (all tcgen05 GEMMs + attention)", "achieved": round(res["value"] / world * w["flop"] / 1e12, 1),
                             "peak": sustained, "unit": "TFLOP/s", "frac": res["step_mfu"], "traffic": None, "peak_source": peak_src + " sustained"},
                "cpu_baseline": None, "clocks": sampler.summary(), "loss": res["loss"]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_ours(args):
    if args.workload != "vcg":
        return run_other_workload(args)
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from src.model.config import MultiModalBartConfig
    from src.model.model import MultiModalBartForConditionalGeneration
    from kmbart.optim import AdamW
    from kmbart import lib as L

    with open(os.path.join(ROOT, "configs", "vcg_base.json")) as f:
        cfg = MultiModalBartConfig.from_dict(json.load(f))
    torch.manual_seed(0)
    model = MultiModalBartForConditionalGeneration(cfg).to(dev)
    model.train()
    step_model = model
    if world > 1:
        model._engine()   # adopt parameters into the flat buffer
        if args.ddp:      # what the reference's scripts do (vcg_train.py:96-98); no overlap with the fused backward
            from torch.nn.parallel import DistributedDataParallel as DDP
            step_model = DDP(model, device_ids=[local_rank], gradient_as_bucket_view=True)
        else:             # all-reduce points inside the backward launch plan, overlapped on NCCL's stream
            from kmbart.parallel import FlatGradReducer
            reducer = FlatGradReducer(model, defer_tail=True)   # AdamW updates the finished regions while the last exchange is in flight
    opt = AdamW(model.parameters(), lr=1e-5)

    dev_batch = make_batch(cfg, 1234 + rank, device=dev)
    host_batch = make_batch(cfg, 1234 + rank, pin=True)

    def step(batch):
        out = step_model.forward(**batch)
        loss = out[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step(dev_batch)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput (inputs already in HBM)
    ms_total = timed(lambda: step(dev_batch), args.steps)
    eng = model._engine()
    launches_step = eng.launches_last + opt.launches_last
    # ---- end to end: pinned host batch -> H2D inside the timed region, loss read back every step
    last = {}

    # kmbart.feed.DeviceFeeder (the repo's replacement of the reference's per-tensor `.to(device)` loop,
    # src/training.py:120-130): every step copies ONE full batch from pinned host memory (the batch of the next
    # step, on a side stream, while this step computes) and reads the loss back; the first batch is staged before
    # the timed region and the last staged batch is left unused, so K steps time exactly K batch copies.
    from kmbart.feed import DeviceFeeder
    feeder = DeviceFeeder(dev, depth=2)
    h2d_seen = []

    def e2e_step():
        b = feeder.get()
        loss = step(b)                               # enqueues forward, backward and the optimizer step
        feeder.release()
        h2d_seen.append(feeder.put(host_batch))      # next batch: copies issued while the GPU works on this one
        last["loss"] = loss.item()

    feeder.put(host_batch)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True
    h2d = sum((sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size())
              for v in host_batch.values())
    assert h2d_seen and h2d_seen[-1] == h2d, "the feeder must move the whole batch every step"

    # ---- dominant kernel: the tcgen05 GEMM.  Timed alone (CUDA events on the launch stream, L2 flushed between launches)
    # in EXACTLY the form the step launches it most expensively: encoder fc1 = bias + exact-erf GELU + bf16 output + bf16
    # pre-activation copy (16 epilogue warps), 12800 x 3072 x 768 — 12 such launches per step (6 encoder layers fwd; the
    # decoder ones have M = 6144).  The step-level figure next to it is step_mfu (whole step vs the sustained peak).
    burst, sustained, hbm, peak_src = load_peaks()
    roof = None
    if rank == 0:
        import ctypes as C
        M, N, K = B_PER_GPU * (R + N_CTX), cfg.encoder_ffn_dim, cfg.d_model
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
        e = L.GemmEpilogue()
        e.alpha, e.act, e.bias = 1.0, L.ACT_GELU, bias.data_ptr()
        e.out_bf16, e.ld_bf16, e.out_preact = out.data_ptr(), N, pre.data_ptr()
        lib = L.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        tot = 0.0
        iters = 20
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.kmb_gemm(A.data_ptr(), W.data_ptr(), M, N, K, K, K, 0, 0, 0, C.byref(e), 0, stream), "gemm")
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                tot += e0.elapsed_time(e1)
        gemm_ms = tot / iters
        ach = 2.0 * M * N * K / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tc05_kernel<256,...,ES=4> encoder fc1 + bias + GELU + pre-activation copy, 12800x3072x768 (as launched in the step)",
                "achieved": round(ach, 1), "peak": burst, "unit": "TFLOP/s", "frac": round(ach / burst, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of exactly these launches inside a training step
                # (ncu over tests/prof_step.py, mean of the 6 encoder fc1 launches: 30.6 MB read + 107.1 MB written;
                # profiles/r02_final_gemm_dram_traffic_per_launch.csv, IDs 3..23); algorithmic operand + output bytes are
                # 181.6 MB, part of the output is still in L2 when the kernel ends
                "traffic": 137645141, "traffic_source": "profiles/r02_final_gemm_dram_traffic_per_launch.csv",
                "us_per_launch": round(gemm_ms * 1e3, 2),
                "peak_source": peak_src + " burst (kernel timed alone, L2 flushed between launches)", "step_mfu": None}
        del A, W, out, pre, flush

    gen = None if args.train_only else bench_generation(model, cfg, dev, rank, world, dist, hbm, peak_src)
    # the path the reference's scripts take unchanged (vcg_train.py:96-98: DistributedDataParallel(find_unused_parameters=True)):
    # torch DDP reduces the .grad views after the fused backward has returned (no overlap) — reported next to the headline
    ddp_leg = None
    if world > 1 and not args.ddp and not args.train_only:
        from torch.nn.parallel import DistributedDataParallel as DDP
        torch.manual_seed(0)
        m2 = MultiModalBartForConditionalGeneration(cfg).to(dev).train()
        m2._engine()
        d2 = DDP(m2, device_ids=[local_rank], find_unused_parameters=True, gradient_as_bucket_view=True)
        o2 = AdamW(m2.parameters(), lr=1e-5)
        b2 = make_batch(cfg, 1234 + rank, device=dev)

        def ddp_step():
            loss = d2.forward(**b2)[0]
            o2.zero_grad()
            loss.backward()
            o2.step()
        for _ in range(3):
            ddp_step()
        ms_ddp = timed(ddp_step, 8)
        ddp_leg = {"grad_exchange": "torch DDP(find_unused_parameters=True), as vcg_train.py:96-98", "ms_per_step": round(ms_ddp / 8, 3),
                   "samples_per_s": round(B_PER_GPU * world * 8 / (ms_ddp * 1e-3), 1)}
        del m2, d2, o2, b2
    extra = {}
    if not args.train_only and args.workload == "vcg":
        del model, opt, dev_batch
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        for name, k in (("pretrain", 6), ("large", 4)):
            try:
                extra[name] = bench_workload(name, dev, rank, world, dist, k, sustained)
            except Exception as ex:   # an extra leg must never take the headline line down
                extra[name] = {"error": repr(ex)[:200]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    sampler.join(timeout=2)
    samples = B_PER_GPU * world * args.steps
    value = samples / (ms_total * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)
    roof["step_mfu"] = round(value / world * FLOP_PER_SAMPLE / 1e12 / sustained, 4)
    roof["step_mfu_peak"] = f"{sustained} TFLOP/s ({peak_src} sustained)"
    cpu = None if args.train_only else cpu_baseline(sample_steps=1)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: KM-BART base VCG fine-tuning step, batch 128/GPU, 36 RoIx2052 + 64 ctx tokens "
                               "(S_e=100), 48 target tokens, dropout 0.1, AdamW lr 1e-5",
                   "global_batch": B_PER_GPU * world, "parallelism": f"dp{world}", "grad_exchange": ("none" if world == 1 else ("torch DDP" if args.ddp else _exchange_text(reducer))),
                   "l2": "per-step working set (~7 GB activations + 1.7 GB optimizer state) far exceeds the 126 MB L2",
                   "reference_script_path": ddp_leg},
        "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3), "api": "kmbart.feed.DeviceFeeder (pinned list-of-tensors batch -> side-stream H2D, one batch per step, overlapped with the previous step) -> model.forward(**batch) + loss.backward() + AdamW.step(), loss.item() each step"},
        "gpu_launches": int(launches_step * args.steps),
        "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(), "loss": last.get("loss"), "gen": gen,
        "workloads": extra,
    }
    if gen is not None:   # the decode step against the HBM roofline, next to the training roofline
        for key in ("rows64_greedy", "rows320_beam5"):
            d = gen["decode_step"][key]
            line["roofline_decode_" + key.split("_")[0]] = {
                "bound": "hbm", "kernel": "persistent decode step + LM head" + (" + greedy select" if "greedy" in key else ""),
                "achieved": d["achieved_GBps"], "peak": hbm, "unit": "GB/s", "frac": d["frac_of_hbm_peak"], "traffic": None,
                "us_per_step": d["us_per_step"], "algorithmic_bytes_per_step": d["algorithmic_bytes_per_step"], "peak_source": peak_src}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


_REF_CFG_KEYS = ["vocab_size", "d_model", "image_feature_size", "encoder_layers", "decoder_layers", "encoder_attention_heads",
                 "decoder_attention_heads", "encoder_ffn_dim", "decoder_ffn_dim", "max_position_embeddings", "dropout", "init_std"]


def _cpu_train_step_fn(batch_size):
    """The reference's CPU path: forward + CE + backward + HF-AdamW in fp32, returns (step_fn, kind).
    kind "reference": the reference's OWN src/model classes (imported from /root/reference through oracle/hf302_shim.py,
    which stands in for the un-installable transformers==3.0.2) — only where that checkout exists (the build container);
    kind "port": the oracle's functional restatement of the same code (the GPU box has no /root/reference)."""
    from oracle import kmbart_oracle as O
    cfg = O.base_config()
    batch = O.synthetic_batch(cfg, batch=batch_size, n_regions=R, n_ctx=N_CTX, tgt_len=S_D, seed=1234)
    if os.path.isdir("/root/reference/src/model") and os.environ.get("KMBART_REFERENCE_ARM", "auto") != "port":
        try:
            from oracle import hf302_shim as S
            mods = S.import_reference()
            torch.manual_seed(0)
            model = mods["model"].MultiModalBartForConditionalGeneration(
                mods["config"].MultiModalBartConfig(**{k: getattr(cfg, k) for k in _REF_CFG_KEYS})).train()
            opt = S.AdamW(model.parameters(), lr=1e-5)      # HF-3.0.2 transformers.AdamW as constructed at vcg_train.py:100

            def ref_step():
                loss = model(**batch)[0]
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss.item()
            return ref_step, "reference"
        except Exception as ex:   # fall back to the port, say why
            print(f"bench.py: reference import failed ({ex!r}); timing the oracle port instead", file=sys.stderr)
    sd = O.init_state_dict(cfg, seed=0)
    names = [k for k in sd if k != "final_logits_bias"]
    for k in names:
        sd[k].requires_grad_(True)
    m = [torch.zeros_like(sd[k]) for k in names]
    v = [torch.zeros_like(sd[k]) for k in names]
    state = {"t": 0}

    def step():
        loss, _, _, _ = O.forward_conditional_generation(sd, cfg, training=True, **batch)
        for k in names:
            sd[k].grad = None
        loss.backward()
        state["t"] += 1
        with torch.no_grad():
            O.adamw_step([sd[k] for k in names], [sd[k].grad for k in names], m, v, state["t"], lr=1e-5)
        return loss.item()
    return step, "port"


def cpu_baseline(sample_steps=1, batch_size=16):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = _cpu_train_step_fn(batch_size)
    step()   # warm-up
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = time.perf_counter() - t0
    what = "the reference's own src/model (through the HF-3.0.2 shim)" if kind == "reference" else "the CPU oracle (port of the reference path)"
    return {"value": round(batch_size * sample_steps / dt, 2), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{sample_steps} fwd+bwd+AdamW step(s) of {what} at batch {batch_size} (same shapes, fp32)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    bs = 16
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = _cpu_train_step_fn(bs)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    dt = time.perf_counter() - t0
    value = bs * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] shapes; each step is a bounded sample (batch 16 of the 128) of the reference's "
                               "CPU path (the reference's own src/model where /root/reference exists, else the oracle restatement of "
                               "src/model + HF-3.0.2 BART), fwd+bwd+AdamW",
                   "global_batch": bs, "parallelism": "cpu"},
        "cpu_baseline": {"value": round(value, 2), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"batch {bs} per step, {args.steps} steps"},
        "e2e": {"value": round(value, 2), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-only", action="store_true", help="experiments: skip the generation legs and the CPU baseline")
    ap.add_argument("--ddp", action="store_true", help="N > 1: wrap in torch DDP (reference scripts' way) instead of FlatGradReducer")
    ap.add_argument("--workload", default="vcg", choices=sorted(WORKLOADS), help="vcg = configs[1] (headline, default; also runs short "
                    "pretrain / large legs); pretrain = configs[2]; large = configs[4] as the headline of the line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
