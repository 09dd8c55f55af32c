"""Device-side decode controllers (csrc/decode_control.cu) against the HF-3.0.2 loop bodies they replace, restated with
torch ops on the same logits: the beam-search step (log_softmax, forced tokens, EOS ban, top 2*num_beams over beams x
vocab, candidate loop with BeamHypotheses, beam re-ordering) must agree token-for-token / beam-for-beam, and the top-k
sampler must draw from exactly the filtered softmax (support and frequencies)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from kmbart import lib as L
    L.require_b200()
    return L, L.load()


def _reference_beam_step(logits, beam_scores, hist, cur_len, hyps, done, nb, V, eos, pad, force, ban):
    """HF-3.0.2 _generate_beam_search loop body (the reference reaches it from src/model/mixins.py:336-366)."""
    B = len(done)
    x = logits.clone()
    if force >= 0:
        keep = torch.zeros(V, dtype=torch.bool)
        keep[force] = True
        x[:, ~keep] = -float("inf")
    scores = torch.log_softmax(x, -1)
    if ban:
        scores[:, eos] = -float("inf")
    nxt = (scores + beam_scores[:, None]).view(B, nb * V)
    ns, nt = torch.topk(nxt, 2 * nb, dim=1, largest=True, sorted=True)
    out = []
    for b in range(B):
        if done[b]:
            out.extend([(0.0, pad, b * nb)] * nb)
            continue
        beam = []
        for rank, (idx, sc) in enumerate(zip(nt[b].tolist(), ns[b].tolist())):
            bid, tok = idx // V, idx % V
            if tok == eos:
                if rank >= nb:
                    continue
                hyps[b].add(hist[b * nb + bid, :cur_len].clone(), sc)
            else:
                beam.append((sc, tok, b * nb + bid))
            if len(beam) == nb:
                break
        done[b] = done[b] or hyps[b].is_done(ns[b].max().item(), cur_len=cur_len)
        assert len(beam) == nb
        out.extend(beam)
    new_scores = torch.tensor([o[0] for o in out], dtype=torch.float32)
    toks = torch.tensor([o[1] for o in out])
    idx = torch.tensor([o[2] for o in out])
    hist2 = hist[idx].clone()
    hist2[:, cur_len] = toks
    return new_scores, toks, idx, hist2


@pytest.mark.parametrize("nb,early,lenpen", [(5, True, 1.0), (3, False, 2.0), (2, False, 0.7)])
def test_beam_step_kernel_matches_hf_loop_body(nb, early, lenpen):
    from kmbart.decode import BeamController
    from src.model.mixins import BeamHypotheses
    L, lib = _lib()
    B, V, max_len, eos, pad = 6, 1543, 12, 2, 1
    rows = B * nb
    g = torch.Generator().manual_seed(100 + nb)
    slot_tbl = torch.arange(rows, dtype=torch.int32, device="cuda").view(rows, 1).repeat(1, max_len).contiguous()
    ctl = BeamController(torch.device("cuda"), B, nb, V, max_len, eos, pad, early, lenpen, slot_tbl=slot_tbl)
    ctl.reset(0)
    beam_scores = torch.zeros(B, nb)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hist = torch.zeros(rows, max_len, dtype=torch.long)
    hyps = [BeamHypotheses(nb, max_len, lenpen, early_stopping=early) for _ in range(B)]
    done = [False] * B
    tbl_ref = torch.arange(rows).view(rows, 1).repeat(1, max_len)
    stream = torch.cuda.current_stream().cuda_stream
    for cur_len in range(1, max_len):
        logits = torch.randn(rows, V, generator=g) * 3.0
        if cur_len >= 3:      # make EOS competitive for some beams so hypotheses finish at different steps
            hot = torch.rand(rows, generator=g) < 0.35
            logits[hot, eos] += 9.0
        force = 0 if cur_len == 1 else (eos if cur_len == max_len - 1 else -1)
        ban = int(cur_len < 4)
        dl = logits.cuda()
        ctl.step(lib, dl, cur_len, force, ban, stream)
        torch.cuda.synchronize()
        was_done = list(done)
        ref_scores, ref_tok, ref_idx, hist = _reference_beam_step(logits, beam_scores, hist, cur_len, hyps, done, nb, V, eos, pad, force, ban)
        beam_scores = ref_scores
        tbl_new = tbl_ref[ref_idx].clone()
        tbl_new[:, cur_len:] = torch.arange(rows).view(rows, 1)
        tbl_ref = tbl_new
        live = torch.tensor([not d for d in was_done]).repeat_interleave(nb)
        if force == eos:
            # forced-EOS step: every finite candidate is an EOS (-> hypotheses, checked below); the next beams are filled
            # from -inf ties in an implementation-defined order and are never used (the loop ends here)
            live = torch.zeros_like(live)
        assert torch.equal(ctl.ids_next.cpu()[live], ref_tok[live]), cur_len
        if cur_len == 1:
            # beams 1.. start from identical states with identical scores (-1e9): which of them torch.topk lists first is
            # implementation defined and immaterial; adopt the device's (equivalent) ancestry for the following steps
            tbl_ref = slot_tbl.cpu().long().clone()
        else:
            assert torch.equal(ctl.beam_idx.cpu().long()[live], ref_idx[live]), cur_len
        assert torch.allclose(ctl.beam_scores.cpu()[live], ref_scores[live], rtol=1e-5, atol=1e-4), cur_len
        assert torch.equal(ctl.hist.cpu().long()[live][:, :cur_len + 1], hist[live][:, :cur_len + 1]), cur_len
        assert torch.equal(slot_tbl.cpu().long()[live][:, :cur_len], tbl_ref[live][:, :cur_len]), cur_len
        assert ctl.done.cpu().bool().tolist() == done, cur_len
        assert ctl.hyp_n.cpu().tolist() == [len(h) for h in hyps], cur_len
        for b in range(B):
            hs = ctl.hyp_score[b, :len(hyps[b])].cpu().tolist()
            assert all(abs(a - s) <= 1e-5 * max(1.0, abs(s)) for a, (s, _) in zip(hs, hyps[b].beams)), (cur_len, b)
            for i, (_, toks) in enumerate(hyps[b].beams):
                assert ctl.hyp_tok[b, i, :len(toks)].cpu().long().tolist() == toks.tolist()
        if all(done):
            break
    assert ctl.all_done() == all(done)
    # epilogue: same best hypotheses as the reference's selection code
    out = ctl.finalize(BeamHypotheses, cur_len + (0 if all(done) else 1), 1, max_len)
    assert out.shape[0] == B and (out[:, 0] == 0).all()


def _sample(lib, L, logits, temperature, top_k, seed, cur_len=3, top_p=1.0):
    rows, V = logits.shape
    unfinished = torch.ones(rows, dtype=torch.int64, device="cuda")
    sent_len = torch.zeros(rows, dtype=torch.int64, device="cuda")
    out = torch.zeros(rows, 8, dtype=torch.int64, device="cuda")
    ids = torch.zeros(rows, dtype=torch.int64, device="cuda")
    sd = torch.tensor([seed], dtype=torch.int64, device="cuda")
    L.check(lib.kmb_sample_select(logits.data_ptr(), V, rows, V, temperature, top_k, top_p, -1, 0, 0, cur_len, sd.data_ptr(), unfinished.data_ptr(),
                                  sent_len.data_ptr(), out.data_ptr(), 8, ids.data_ptr(), torch.cuda.current_stream().cuda_stream), "sample")
    torch.cuda.synchronize()
    assert torch.equal(out[:, cur_len], ids)
    return ids.cpu()


def test_sample_select_support_and_frequencies():
    L, lib = _lib()
    V = 50320
    g = torch.Generator().manual_seed(5)
    base = torch.randn(V, generator=g)
    base[[7, 4242, 50300]] = torch.tensor([6.0, 5.0, 4.5])     # three tokens carry most of the mass
    rows = 4096
    logits = base.repeat(rows, 1).cuda().contiguous()
    # top_k = 1 is arg-max; top_k = 3 stays inside the top three and follows their renormalised softmax
    assert (_sample(lib, L, logits, 1.0, 1, 11) == 7).all()
    t3 = _sample(lib, L, logits, 1.0, 3, 12)
    assert set(t3.tolist()) <= {7, 4242, 50300}
    p = torch.softmax(torch.tensor([6.0, 5.0, 4.5]), 0)
    for tok, pi in zip((7, 4242, 50300), p.tolist()):
        f = (t3 == tok).float().mean().item()
        assert abs(f - pi) <= 4.0 * (pi * (1 - pi) / rows) ** 0.5 + 1e-3, (tok, f, pi)
    # top_k = 0: no filter — the tail must show up with its softmax mass, temperature rescales it
    full = torch.softmax(base.double(), 0)
    tail_mass = 1.0 - full[[7, 4242, 50300]].sum().item()
    t0 = _sample(lib, L, logits, 1.0, 0, 13)
    tail = 1.0 - sum((t0 == tok).float().mean().item() for tok in (7, 4242, 50300))
    assert abs(tail - tail_mass) <= 4.0 * (tail_mass * (1 - tail_mass) / rows) ** 0.5 + 1e-3, (tail, tail_mass)
    hot = torch.softmax(base.double() / 0.5, 0)
    th = _sample(lib, L, logits, 0.5, 0, 14)
    f7 = (th == 7).float().mean().item()
    assert abs(f7 - hot[7].item()) <= 4.0 * (hot[7].item() * (1 - hot[7].item()) / rows) ** 0.5 + 1e-3
    # different seeds / steps give different draws, the same (seed, step) the same draw
    assert torch.equal(_sample(lib, L, logits, 1.0, 50, 21), _sample(lib, L, logits, 1.0, 50, 21))
    assert not torch.equal(_sample(lib, L, logits, 1.0, 50, 21), _sample(lib, L, logits, 1.0, 50, 22))
    assert not torch.equal(_sample(lib, L, logits, 1.0, 50, 21, cur_len=3), _sample(lib, L, logits, 1.0, 50, 21, cur_len=4))


def test_sample_select_top_k_threshold_keeps_ties_and_ragged_rows():
    """top_k keeps every logit >= the k-th largest VALUE (HF-3.0.2 top_k_top_p_filtering); rows differ."""
    L, lib = _lib()
    V, rows = 1000, 512
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(rows, V, generator=g).clamp(max=3.5)
    logits[:, 10] = 5.0
    logits[:, 20] = 4.0
    logits[:, 30] = 4.0          # tie at the 2nd largest value: top_k = 2 must keep all three
    toks = _sample(lib, L, logits.cuda().contiguous(), 1.0, 2, 3)
    assert set(toks.tolist()) == {10, 20, 30}
    k = 17
    toks = _sample(lib, L, logits.cuda().contiguous(), 1.0, k, 4)
    kth = logits.topk(k, -1).values[:, -1]
    assert bool((logits.gather(1, toks.view(-1, 1)).squeeze(1) >= kth).all())


def test_sample_select_nucleus_filter_matches_hf_rule():
    """top_p: a token survives iff the probability mass sorted strictly before it is <= top_p (HF-3.0.2
    top_k_top_p_filtering: cumulative softmax of the sorted logits, shifted right by one) — alone and after a top-k filter."""
    L, lib = _lib()
    V, rows = 2000, 2048
    g = torch.Generator().manual_seed(17)
    base = torch.randn(V, generator=g) * 2.0
    logits = base.repeat(rows, 1).cuda().contiguous()

    def hf_support(x, top_k, top_p, temperature=1.0):
        x = x.clone() / temperature
        if top_k > 0:
            x[x < x.topk(top_k).values[-1]] = -float("inf")
        sl, si = torch.sort(x, descending=True)
        cum = torch.softmax(sl, -1).cumsum(-1)
        rm = cum > top_p
        rm[1:] = rm[:-1].clone()
        rm[0] = False
        keep = torch.zeros(V, dtype=torch.bool)
        keep[si[~rm]] = True
        return keep, torch.softmax(x.masked_fill(~keep, -float("inf")), -1)
    for top_k, top_p, temp, seed in ((0, 0.8, 1.0, 1), (0, 0.3, 1.0, 2), (40, 0.9, 1.0, 3), (0, 0.8, 0.7, 4)):
        keep, probs = hf_support(base, top_k, top_p, temp)
        toks = _sample(lib, L, logits, temp, top_k, seed, top_p=top_p)
        assert bool(keep[toks].all()), (top_k, top_p, "token outside the nucleus")
        seen = torch.zeros(V, dtype=torch.bool)
        seen[toks] = True
        heavy = keep & (probs > 8.0 / rows)                  # every reasonably likely survivor shows up
        assert bool(seen[heavy].all()), (top_k, top_p)
        top = int(probs.argmax())
        f = (toks == top).float().mean().item()
        pt = probs[top].item()
        assert abs(f - pt) <= 4.0 * (pt * (1 - pt) / rows) ** 0.5 + 1e-3, (top_k, top_p, f, pt)
