"""GenerationMixin and FromPretrainedMixin — drop-in for the reference's src/model/mixins.py.

generate() keeps the reference's signature, argument checks and control flow
(src/model/mixins.py:33-384): encoder once, beams/returns expanded, then the greedy / sampling /
beam-search loops of HF-3.0.2 generation_utils that the reference inherits.  Every model call
inside those loops is the KV-cached sm_100a decode step (kmbart.engine.Engine.decoder_step);
prepare_inputs_for_generation / adjust_logits_during_generation / _reorder_cache keep their
reference semantics (:386-434) and cache structure.

from_pretrained()/save_pretrained() read/write the reference's checkpoint format
(config.json + pytorch_model.bin, src/model/mixins.py:551-883) including `partial_load`
slice-copy for shape-mismatched tensors (:511-528); only local paths are supported (no network)."""
import logging
import os
from typing import Iterable, Optional

import torch
from torch.nn import functional as F

logger = logging.getLogger(__name__)
WEIGHTS_NAME = "pytorch_model.bin"


def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("Inf"), min_tokens_to_keep=1):
    """HF-3.0.2 generation_utils.top_k_top_p_filtering semantics (in place)."""
    if top_k > 0:
        top_k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
        logits[remove] = filter_value
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        rm = cum > top_p
        if min_tokens_to_keep > 1:
            rm[..., :min_tokens_to_keep] = 0
        rm[..., 1:] = rm[..., :-1].clone()
        rm[..., 0] = 0
        logits[rm.scatter(1, sorted_idx, rm)] = filter_value
    return logits


class BeamHypotheses(object):
    """n-best list of finished hypotheses for one batch element (HF-3.0.2 semantics)."""

    def __init__(self, num_beams, max_length, length_penalty, early_stopping):
        self.max_length = max_length - 1
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                ranked = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


class GenerationMixin:
    @torch.no_grad()
    def generate(
            self,
            input_ids: Optional[torch.LongTensor] = None,
            image_features=None,
            max_length: Optional[int] = None,
            min_length: Optional[int] = None,
            do_sample: Optional[bool] = None,
            early_stopping: Optional[bool] = None,
            num_beams: Optional[int] = None,
            temperature: Optional[float] = None,
            top_k: Optional[int] = None,
            top_p: Optional[float] = None,
            repetition_penalty: Optional[float] = None,
            bad_words_ids: Optional[Iterable[int]] = None,
            bos_token_id: Optional[int] = None,
            pad_token_id: Optional[int] = None,
            eos_token_id: Optional[int] = None,
            length_penalty: Optional[float] = None,
            no_repeat_ngram_size: Optional[int] = None,
            num_return_sequences: Optional[int] = None,
            attention_mask: Optional[torch.LongTensor] = None,
            decoder_start_token_id: Optional[int] = None,
            use_cache: Optional[bool] = None,
            **model_specific_kwargs
    ) -> torch.LongTensor:
        cfg = self.config
        pick = lambda v, name: v if v is not None else getattr(cfg, name)
        max_length, min_length = pick(max_length, "max_length"), pick(min_length, "min_length")
        do_sample, early_stopping = pick(do_sample, "do_sample"), pick(early_stopping, "early_stopping")
        use_cache, num_beams = pick(use_cache, "use_cache"), pick(num_beams, "num_beams")
        temperature, top_k, top_p = pick(temperature, "temperature"), pick(top_k, "top_k"), pick(top_p, "top_p")
        repetition_penalty = pick(repetition_penalty, "repetition_penalty")
        bos_token_id, pad_token_id = pick(bos_token_id, "bos_token_id"), pick(pad_token_id, "pad_token_id")
        eos_token_id, length_penalty = pick(eos_token_id, "eos_token_id"), pick(length_penalty, "length_penalty")
        no_repeat_ngram_size = pick(no_repeat_ngram_size, "no_repeat_ngram_size")
        bad_words_ids = pick(bad_words_ids, "bad_words_ids")
        num_return_sequences = pick(num_return_sequences, "num_return_sequences")
        decoder_start_token_id = pick(decoder_start_token_id, "decoder_start_token_id")

        batch_size = input_ids.shape[0] if input_ids is not None else 1

        assert isinstance(max_length, int) and max_length > 0, "max_length should be a strictly positive integer."
        assert isinstance(min_length, int) and min_length >= 0, "min_length should be a positive integer."
        assert isinstance(do_sample, bool), "do_sample should be a boolean."
        assert isinstance(early_stopping, bool), "early_stopping should be a boolean."
        assert isinstance(use_cache, bool), "use_cache should be a boolean."
        assert isinstance(num_beams, int) and num_beams > 0, "num_beams should be a strictly positive integer."
        assert temperature > 0, "temperature should be strictly positive."
        assert isinstance(top_k, int) and top_k >= 0, "top_k should be a positive integer."
        assert 0 <= top_p <= 1, "top_p should be between 0 and 1."
        assert repetition_penalty >= 1.0, "repetition_penalty should be >= 1."
        assert input_ids is not None or (isinstance(bos_token_id, int) and bos_token_id >= 0), \
            "If input_ids is not defined, bos_token_id should be a positive integer."
        assert pad_token_id is None or (isinstance(pad_token_id, int) and (pad_token_id >= 0)), \
            "pad_token_id should be a positive integer."
        assert (eos_token_id is None) or (isinstance(eos_token_id, int) and (eos_token_id >= 0)), \
            "eos_token_id should be a positive integer."
        assert length_penalty > 0, "length_penalty should be strictly positive."
        assert isinstance(no_repeat_ngram_size, int) and no_repeat_ngram_size >= 0, \
            "no_repeat_ngram_size should be a positive integer."
        assert isinstance(num_return_sequences, int) and num_return_sequences > 0, \
            "num_return_sequences should be a strictly positive integer."
        assert bad_words_ids is None or isinstance(bad_words_ids, list) and isinstance(bad_words_ids[0], list), \
            "bad_words_ids is either None or a list of lists of tokens that should not be generated"

        device = next(self.parameters()).device
        if input_ids is None:
            input_ids = torch.full((batch_size, 1), bos_token_id, dtype=torch.long, device=device)
        else:
            assert input_ids.dim() == 2, "Input prompt should be of shape (batch_size, sequence length)."

        if do_sample is False:
            if num_beams == 1:
                assert num_return_sequences == 1, \
                    "Greedy decoding will always produce the same output for num_beams == 1 and " \
                    "num_return_sequences > 1. Please set num_return_sequences = 1"
            else:
                assert num_beams >= num_return_sequences, \
                    "Greedy beam search decoding cannot return more sequences than it has beams. " \
                    "Please set num_beams >= num_return_sequences"

        if (attention_mask is None) and (pad_token_id is not None) and (pad_token_id in input_ids):
            attention_mask = input_ids.ne(pad_token_id).long()
        elif attention_mask is None:
            attention_mask = input_ids.new_ones(input_ids.shape)

        if pad_token_id is None and eos_token_id is not None:
            logger.warning("Setting pad_token_id to {} (first eos_token_id) to generate sequence".format(eos_token_id))
            pad_token_id = eos_token_id

        vocab_size = cfg.vocab_size
        if do_sample:
            effective_batch_size, effective_batch_mult = batch_size * num_return_sequences, num_return_sequences
        else:
            effective_batch_size, effective_batch_mult = batch_size, 1

        assert cfg.is_encoder_decoder
        if decoder_start_token_id is None:
            decoder_start_token_id = bos_token_id
        assert decoder_start_token_id is not None, \
            "decoder_start_token_id or bos_token_id has to be defined for encoder-decoder generation"
        encoder = self.get_encoder()
        encoder_outputs: tuple = encoder(input_ids, image_features=image_features, attention_mask=attention_mask)

        # ---- fast path: preallocated caches, one CUDA graph per step, device-side token selection (kmbart/decode.py).
        # Same token semantics as the legacy loops below; knobs the decode chain does not implement fall through.
        fast_ok = (use_cache and repetition_penalty == 1.0 and no_repeat_ngram_size == 0 and bad_words_ids is None
                   and not model_specific_kwargs and max_length <= 256 and input_ids.shape[1] <= 256
                   and (num_beams == 1 or (not do_sample and max_length > 2))
                   and getattr(self, "_fast_generate", True) and self.precision != "fp32" and self._select_kernels_fit(num_beams))
        if fast_ok:
            return self._generate_fast(encoder_outputs[0], attention_mask, batch_size, effective_batch_size, effective_batch_mult,
                                       num_beams, max_length, min_length, do_sample, early_stopping, temperature, top_k,
                                       pad_token_id, eos_token_id, length_penalty, num_return_sequences, decoder_start_token_id,
                                       vocab_size, top_p=top_p)

        if num_return_sequences > 1 or num_beams > 1:
            input_ids_len = input_ids.shape[-1]
            attention_mask = attention_mask.unsqueeze(1).expand(batch_size, effective_batch_mult * num_beams, input_ids_len)
            attention_mask = attention_mask.contiguous().view(effective_batch_size * num_beams, input_ids_len)

        input_ids = torch.full((effective_batch_size * num_beams, 1), decoder_start_token_id, dtype=torch.long, device=device)
        cur_len = 1
        assert batch_size == encoder_outputs[0].shape[0], \
            f"expected encoder_outputs[0] to have 1st dimension bs={batch_size}, got {encoder_outputs[0].shape[0]} "
        expanded_batch_idxs = (
            torch.arange(batch_size).view(-1, 1).repeat(1, num_beams * effective_batch_mult).view(-1).to(input_ids.device)
        )
        encoder_outputs = (encoder_outputs[0].index_select(0, expanded_batch_idxs), *encoder_outputs[1:])

        assert cur_len < max_length, \
            f"The context has {cur_len} number of tokens, but max_length is only {max_length}. " \
            f"Please make sure that max_length is bigger than the number of tokens, " \
            f"by setting either generate(max_length=...,...) or config.max_length = ..."

        common = dict(cur_len=cur_len, max_length=max_length, min_length=min_length, do_sample=do_sample,
                      temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty,
                      no_repeat_ngram_size=no_repeat_ngram_size, bad_words_ids=bad_words_ids, pad_token_id=pad_token_id,
                      eos_token_id=eos_token_id, batch_size=effective_batch_size, encoder_outputs=encoder_outputs,
                      attention_mask=attention_mask, use_cache=use_cache, model_specific_kwargs=model_specific_kwargs)
        if num_beams > 1:
            return self._generate_beam_search(input_ids, early_stopping=early_stopping,
                                              num_return_sequences=num_return_sequences, length_penalty=length_penalty,
                                              num_beams=num_beams, vocab_size=vocab_size, **common)
        return self._generate_no_beam_search(input_ids, **common)


    def _select_kernels_fit(self, num_beams):
        """the sampling / beam controllers keep one logits row in a CTA's shared memory"""
        from kmbart import lib as L
        return self.config.vocab_size <= L.load().kmb_select_max_vocab() and 2 * num_beams <= self.config.vocab_size and num_beams <= 64

    # ------------------------------------------------------------------ fast decode (same semantics, device-side loop)
    def _generate_fast(self, enc_hidden, attention_mask, batch_size, effective_batch_size, effective_batch_mult, num_beams,
                       max_length, min_length, do_sample, early_stopping, temperature, top_k, pad_token_id, eos_token_id,
                       length_penalty, num_return_sequences, decoder_start_token_id, vocab_size, top_p=1.0):
        from kmbart.decode import get_session
        cfg = self.config
        eng = self._engine()
        eng.sync_shadow()
        B, Se = enc_hidden.shape[0], enc_hidden.shape[1]
        rows = effective_batch_size * num_beams
        has_pad = bool((attention_mask == 0).any().item()) if attention_mask is not None else False
        sess = get_session(eng, B, Se, rows, max_length, has_pad)
        sess.begin(enc_hidden, attention_mask, decoder_start_token_id, use_tbl=num_beams > 1)
        flb = self.final_logits_bias
        if num_beams == 1:
            sel = dict(do_sample=bool(do_sample), temperature=float(temperature), top_k=int(top_k), top_p=float(top_p), eos=eos_token_id,
                       pad=pad_token_id, min_length=int(min_length))
            steps = max_length - 1
            for t in range(steps):
                sess.step(t, flb, sel)
                if eos_token_id is not None and (t % 8 == 7) and t + 1 < steps and int(sess.unfinished.max().item()) == 0:
                    break
            sent_len = sess.sent_len.clone()
            width = int(sent_len.max().item())
            decoded = sess.out[:, :width].clone()
            if int(sent_len.min().item()) != width:
                assert pad_token_id is not None, "`Pad_token_id` has to be defined if batches have different lengths"
                beyond = torch.arange(width, device=decoded.device).unsqueeze(0) >= sent_len.unsqueeze(1)
                decoded[beyond] = pad_token_id
            return decoded
        # ---- beam search: model step + LM head + device-side beam controller in one graph per token; the host only polls
        # "every batch element done" every fourth step and finalises the hypotheses once at the end
        ctl = sess.beam_controller(num_beams, early_stopping, length_penalty, eos_token_id, pad_token_id)
        ctl.reset(decoder_start_token_id)
        steps = max_length - 1
        cur_len = 1
        for t in range(steps):
            cur_len = t + 1
            # adjust_logits_during_generation (src/model/mixins.py:386-402): BOS forced at cur_len 1, EOS at max_length - 1
            force = cfg.bos_token_id if cur_len == 1 else (cfg.eos_token_id if (cur_len == max_length - 1 and cfg.eos_token_id is not None) else -1)
            ban = int(eos_token_id is not None and cur_len < min_length)
            sess.step(t, flb, None, beam=(ctl, int(force), ban))
            cur_len = t + 2
            if eos_token_id is not None and (t % 4 == 3) and t + 1 < steps and ctl.all_done():
                break
        return ctl.finalize(BeamHypotheses, cur_len, num_return_sequences, max_length)

    # ------------------------------------------------------------------ loops (HF-3.0.2 generation_utils semantics)
    def _use_cache(self, outputs, use_cache):
        return not (len(outputs) <= 1 or use_cache is False)

    def enforce_repetition_penalty_(self, lprobs, batch_size, num_beams, prev_output_tokens, repetition_penalty):
        for i in range(batch_size * num_beams):
            for previous_token in set(prev_output_tokens[i].tolist()):
                if lprobs[i, previous_token] < 0:
                    lprobs[i, previous_token] *= repetition_penalty
                else:
                    lprobs[i, previous_token] /= repetition_penalty

    def postprocess_next_token_scores(self, scores, input_ids, no_repeat_ngram_size, bad_words_ids, cur_len, min_length,
                                      max_length, eos_token_id, repetition_penalty, batch_size, num_beams):
        if repetition_penalty != 1.0:
            self.enforce_repetition_penalty_(scores, batch_size, num_beams, input_ids, repetition_penalty)
        if eos_token_id is not None and cur_len < min_length:
            scores[:, eos_token_id] = -float("inf")
        if no_repeat_ngram_size > 0:
            n = no_repeat_ngram_size
            for i, seq in enumerate(input_ids.tolist()):
                if cur_len + 1 < n:
                    continue
                prefix = tuple(seq[cur_len + 1 - n:cur_len])
                banned = [seq[j + n - 1] for j in range(len(seq) - n + 1) if tuple(seq[j:j + n - 1]) == prefix]
                if banned:
                    scores[i, banned] = -float("inf")
        if bad_words_ids is not None:
            for i, seq in enumerate(input_ids.tolist()):
                for bw in bad_words_ids:
                    if len(bw) == 1 or (len(bw) - 1 <= len(seq) and seq[-(len(bw) - 1):] == bw[:-1]):
                        scores[i, bw[-1]] = -float("inf")
        return scores

    def _generate_no_beam_search(self, input_ids, cur_len, max_length, min_length, do_sample, temperature, top_k, top_p,
                                 repetition_penalty, no_repeat_ngram_size, bad_words_ids, pad_token_id, eos_token_id,
                                 batch_size, encoder_outputs, attention_mask, use_cache, model_specific_kwargs):
        unfinished_sents = input_ids.new(batch_size).fill_(1)
        sent_lengths = input_ids.new(batch_size).fill_(max_length)
        past = (encoder_outputs, None) if encoder_outputs is not None else None
        while cur_len < max_length:
            model_inputs = self.prepare_inputs_for_generation(input_ids, past=past, attention_mask=attention_mask,
                                                              use_cache=use_cache, **model_specific_kwargs)
            outputs = self(**model_inputs)
            next_token_logits = outputs[0][:, -1, :]
            scores = self.postprocess_next_token_scores(
                scores=next_token_logits, input_ids=input_ids, no_repeat_ngram_size=no_repeat_ngram_size,
                bad_words_ids=bad_words_ids, cur_len=cur_len, min_length=min_length, max_length=max_length,
                eos_token_id=eos_token_id, repetition_penalty=repetition_penalty, batch_size=batch_size, num_beams=1)
            if self._use_cache(outputs, use_cache):
                past = outputs[1]
            if do_sample:
                if temperature != 1.0:
                    scores = scores / temperature
                next_token_logscores = top_k_top_p_filtering(scores, top_k=top_k, top_p=top_p)
                probs = F.softmax(next_token_logscores, dim=-1)
                next_token = torch.multinomial(probs, num_samples=1).squeeze(1)
            else:
                next_token = torch.argmax(next_token_logits, dim=-1)
            if eos_token_id is not None:
                tokens_to_add = next_token * unfinished_sents + (pad_token_id) * (1 - unfinished_sents)
            else:
                tokens_to_add = next_token
            input_ids = torch.cat([input_ids, tokens_to_add.unsqueeze(-1)], dim=-1)
            cur_len = cur_len + 1
            if eos_token_id is not None:
                eos_in_sents = tokens_to_add == eos_token_id
                newly_done = unfinished_sents.mul(eos_in_sents.long()).bool()
                sent_lengths.masked_fill_(newly_done, cur_len)
                unfinished_sents.mul_((~eos_in_sents).long())
            if unfinished_sents.max() == 0:
                break
        if sent_lengths.min().item() != sent_lengths.max().item():
            assert pad_token_id is not None, "`Pad_token_id` has to be defined if batches have different lengths"
            decoded = input_ids.new(batch_size, sent_lengths.max().item()).fill_(pad_token_id)
        else:
            decoded = input_ids
        for hypo_idx, hypo in enumerate(input_ids):
            decoded[hypo_idx, : sent_lengths[hypo_idx]] = hypo[: sent_lengths[hypo_idx]]
        return decoded

    def _generate_beam_search(self, input_ids, cur_len, max_length, min_length, do_sample, early_stopping, temperature,
                              top_k, top_p, repetition_penalty, no_repeat_ngram_size, bad_words_ids, pad_token_id,
                              eos_token_id, batch_size, num_return_sequences, length_penalty, num_beams, vocab_size,
                              encoder_outputs, attention_mask, use_cache, model_specific_kwargs):
        if (not do_sample and repetition_penalty == 1.0 and no_repeat_ngram_size == 0 and bad_words_ids is None
                and max_length > 2 and cur_len == 1 and input_ids.is_cuda and getattr(self, "_device_controller", True)
                and self._select_kernels_fit(num_beams)):
            return self._beam_search_device_controller(input_ids, max_length, min_length, early_stopping, pad_token_id, eos_token_id,
                                                       batch_size, num_return_sequences, length_penalty, num_beams, vocab_size,
                                                       encoder_outputs, attention_mask, use_cache, model_specific_kwargs)
        generated_hyps = [BeamHypotheses(num_beams, max_length, length_penalty, early_stopping=early_stopping)
                          for _ in range(batch_size)]
        beam_scores = torch.zeros((batch_size, num_beams), dtype=torch.float, device=input_ids.device)
        if do_sample is False:
            beam_scores[:, 1:] = -1e9
        beam_scores = beam_scores.view(-1)
        past = (encoder_outputs, None) if encoder_outputs is not None else None
        done = [False for _ in range(batch_size)]
        while cur_len < max_length:
            model_inputs = self.prepare_inputs_for_generation(input_ids, past=past, attention_mask=attention_mask,
                                                              use_cache=use_cache, **model_specific_kwargs)
            outputs = self(**model_inputs)
            next_token_logits = outputs[0][:, -1, :]
            if self._use_cache(outputs, use_cache):
                past = outputs[1]
            if self.config.is_encoder_decoder and do_sample is False:
                next_token_logits = self.adjust_logits_during_generation(next_token_logits, cur_len=cur_len,
                                                                         max_length=max_length)
            scores = F.log_softmax(next_token_logits, dim=-1)
            scores = self.postprocess_next_token_scores(
                scores=scores, input_ids=input_ids, no_repeat_ngram_size=no_repeat_ngram_size, bad_words_ids=bad_words_ids,
                cur_len=cur_len, min_length=min_length, max_length=max_length, eos_token_id=eos_token_id,
                repetition_penalty=repetition_penalty, batch_size=batch_size, num_beams=num_beams)
            assert scores.shape == (batch_size * num_beams, vocab_size)
            if do_sample:
                _scores = scores + beam_scores[:, None].expand_as(scores)
                if temperature != 1.0:
                    _scores = _scores / temperature
                _scores = top_k_top_p_filtering(_scores, top_k=top_k, top_p=top_p, min_tokens_to_keep=2)
                _scores = _scores.contiguous().view(batch_size, num_beams * vocab_size)
                probs = F.softmax(_scores, dim=-1)
                next_tokens = torch.multinomial(probs, num_samples=2 * num_beams)
                next_scores = torch.gather(_scores, -1, next_tokens)
                next_scores, next_scores_indices = torch.sort(next_scores, descending=True, dim=1)
                next_tokens = torch.gather(next_tokens, -1, next_scores_indices)
            else:
                next_scores = scores + beam_scores[:, None].expand_as(scores)
                next_scores = next_scores.view(batch_size, num_beams * vocab_size)
                next_scores, next_tokens = torch.topk(next_scores, 2 * num_beams, dim=1, largest=True, sorted=True)
            assert next_scores.size() == next_tokens.size() == (batch_size, 2 * num_beams)

            # one device->host transfer per step instead of the reference's per-candidate .item()
            tok_host, score_host = next_tokens.tolist(), next_scores.tolist()
            next_batch_beam = []
            for batch_idx in range(batch_size):
                if done[batch_idx]:
                    assert len(generated_hyps[batch_idx]) >= num_beams
                    assert eos_token_id is not None and pad_token_id is not None
                    next_batch_beam.extend([(0, pad_token_id, 0)] * num_beams)
                    continue
                next_sent_beam = []
                for beam_token_rank, (beam_token_id, beam_token_score) in enumerate(zip(tok_host[batch_idx], score_host[batch_idx])):
                    beam_id = beam_token_id // vocab_size
                    token_id = beam_token_id % vocab_size
                    effective_beam_id = batch_idx * num_beams + beam_id
                    if (eos_token_id is not None) and (token_id == eos_token_id):
                        if beam_token_rank >= num_beams:
                            continue
                        generated_hyps[batch_idx].add(input_ids[effective_beam_id].clone(), beam_token_score)
                    else:
                        next_sent_beam.append((beam_token_score, token_id, effective_beam_id))
                    if len(next_sent_beam) == num_beams:
                        break
                done[batch_idx] = done[batch_idx] or generated_hyps[batch_idx].is_done(max(score_host[batch_idx]), cur_len=cur_len)
                assert len(next_sent_beam) == num_beams, "Beam should always be full"
                next_batch_beam.extend(next_sent_beam)
            if all(done):
                break
            beam_scores = beam_scores.new([x[0] for x in next_batch_beam])
            beam_tokens = input_ids.new([x[1] for x in next_batch_beam])
            beam_idx = input_ids.new([x[2] for x in next_batch_beam])
            input_ids = input_ids[beam_idx, :]
            input_ids = torch.cat([input_ids, beam_tokens.unsqueeze(1)], dim=-1)
            cur_len = cur_len + 1
            if past is not None:
                past = self._reorder_cache(past, beam_idx)

        for batch_idx in range(batch_size):
            if done[batch_idx]:
                continue
            for beam_id in range(num_beams):
                effective_beam_id = batch_idx * num_beams + beam_id
                generated_hyps[batch_idx].add(input_ids[effective_beam_id], beam_scores[effective_beam_id].item())

        output_batch_size = batch_size if do_sample else batch_size * num_return_sequences
        per_batch = 1 if do_sample else num_return_sequences
        sent_lengths = input_ids.new(output_batch_size)
        best = []
        for i, hypotheses in enumerate(generated_hyps):
            sorted_hyps = sorted(hypotheses.beams, key=lambda x: x[0])
            for j in range(per_batch):
                best_hyp = sorted_hyps.pop()[1]
                sent_lengths[per_batch * i + j] = len(best_hyp)
                best.append(best_hyp)
        if sent_lengths.min().item() != sent_lengths.max().item():
            assert pad_token_id is not None, "`Pad_token_id` has to be defined"
            sent_max_len = min(sent_lengths.max().item() + 1, max_length)
            decoded = input_ids.new(output_batch_size, sent_max_len).fill_(pad_token_id)
            for i, hypo in enumerate(best):
                decoded[i, : sent_lengths[i]] = hypo
                if sent_lengths[i] < max_length:
                    decoded[i, sent_lengths[i]] = eos_token_id
        else:
            decoded = torch.stack(best).type(torch.long).to(next(self.parameters()).device)
        return decoded

    def _beam_search_device_controller(self, input_ids, max_length, min_length, early_stopping, pad_token_id, eos_token_id, batch_size,
                                       num_return_sequences, length_penalty, num_beams, vocab_size, encoder_outputs, attention_mask,
                                       use_cache, model_specific_kwargs):
        """_generate_beam_search with the model called eagerly (legacy cache dicts, or the fp32 parity mode) and the loop body
        — log_softmax, forced tokens, top 2 * num_beams, hypothesis bookkeeping, beam re-ordering — on the device
        (kmbart.decode.BeamController): the only host synchronisation is an "all done" poll every fourth step."""
        from kmbart.decode import BeamController
        from kmbart import lib as L
        cfg = self.config
        dev = input_ids.device
        key = (batch_size, num_beams, max_length, eos_token_id, pad_token_id, bool(early_stopping), float(length_penalty), dev)
        cache = self.__dict__.setdefault("_beam_controllers", {})
        ctl = cache.get(key)
        if ctl is None:
            cache.clear()      # one resident controller is enough for this path
            ctl = cache[key] = BeamController(dev, batch_size, num_beams, vocab_size, max_length, eos_token_id, pad_token_id,
                                              early_stopping, length_penalty)
        ctl.reset(int(input_ids[0, 0].item()))
        lib = L.load()
        past = (encoder_outputs, None) if encoder_outputs is not None else None
        cur_len = 1
        while cur_len < max_length:
            dec_ids = ctl.hist[:, :cur_len].long()
            model_inputs = self.prepare_inputs_for_generation(dec_ids, past=past, attention_mask=attention_mask, use_cache=use_cache,
                                                              **model_specific_kwargs)
            outputs = self(**model_inputs)
            logits = outputs[0][:, -1, :].float().contiguous()
            if self._use_cache(outputs, use_cache):
                past = outputs[1]
            force = cfg.bos_token_id if cur_len == 1 else (cfg.eos_token_id if (cur_len == max_length - 1 and cfg.eos_token_id is not None) else -1)
            ban = int(eos_token_id is not None and cur_len < min_length)
            ctl.step(lib, logits, cur_len, int(force), ban, torch.cuda.current_stream(dev).cuda_stream)
            step_no = cur_len
            cur_len += 1
            if eos_token_id is not None and step_no % 4 == 0 and cur_len < max_length and ctl.all_done():
                break
            if past is not None and past[1] is not None:
                past = self._reorder_cache(past, ctl.beam_idx.long())
        return ctl.finalize(BeamHypotheses, cur_len, num_return_sequences, max_length)

    # ------------------------------------------------------------------ reference hooks (src/model/mixins.py:386-434)
    def prepare_inputs_for_generation(self, decoder_input_ids, past, attention_mask, use_cache, **kwargs):
        assert past is not None, "past has to be defined for encoder_outputs"
        encoder_outputs, decoder_cached_states = past
        return {
            "input_ids": None,
            "image_features": None,
            "encoder_outputs": encoder_outputs,
            "decoder_cached_states": decoder_cached_states,
            "decoder_input_ids": decoder_input_ids,
            "attention_mask": attention_mask,
            "use_cache": use_cache,
        }

    def adjust_logits_during_generation(self, logits, cur_len, max_length):
        if cur_len == 1:
            self._force_token_ids_generation(logits, self.config.bos_token_id)
        if cur_len == max_length - 1 and self.config.eos_token_id is not None:
            self._force_token_ids_generation(logits, self.config.eos_token_id)
        return logits

    def _force_token_ids_generation(self, scores, token_ids) -> None:
        """Everything except token_ids gets -inf (the reference builds a 50k-element Python list per
        call, src/model/mixins.py:407-417; same result with one masked fill)."""
        if isinstance(token_ids, int):
            token_ids = [token_ids]
        assert len(scores.shape) == 2, "scores should be of rank 2 with shape: [batch_size, vocab_size]"
        keep = torch.zeros(scores.shape[1], dtype=torch.bool, device=scores.device)
        keep[token_ids] = True
        scores.masked_fill_(~keep, -float("inf"))

    @staticmethod
    def _reorder_cache(past, beam_idx):
        ((enc_out, enc_mask), decoder_cached_states) = past
        reordered_past = []
        for layer_past in decoder_cached_states:
            layer_past_new = {}
            for attn_key, attn_cache in layer_past.items():
                layer_past_new[attn_key] = {k: (v.index_select(0, beam_idx) if v is not None else None)
                                            for k, v in attn_cache.items()}
            reordered_past.append(layer_past_new)
        new_enc_out = enc_out if enc_out is None else enc_out.index_select(0, beam_idx)
        new_enc_mask = enc_mask if enc_mask is None else enc_mask.index_select(0, beam_idx)
        return ((new_enc_out, new_enc_mask), reordered_past)

    def get_encoder(self):
        return self.model.encoder

    def get_output_embeddings(self):
        return self.model.get_output_embeddings()

    def resize_token_embeddings(self, new_num_tokens: int):
        raise NotImplementedError("resize_token_embeddings is not used by the KM-BART scripts; vocab is fixed at "
                                  "config.vocab_size (50320) and partial_load handles bart-base checkpoints")


class FromPretrainedMixin:
    """Checkpoint IO with the reference's semantics (src/model/mixins.py:458-883)."""

    def save_pretrained(self, save_directory):
        assert os.path.isdir(save_directory), "Saving path should be a directory where the model and configuration can be saved"
        model_to_save = self.module if hasattr(self, "module") else self
        model_to_save.config.architectures = [model_to_save.__class__.__name__]
        model_to_save.config.save_pretrained(save_directory)
        output_model_file = os.path.join(save_directory, WEIGHTS_NAME)
        state = {k: v.detach().to("cpu").clone() for k, v in model_to_save.state_dict().items()}
        torch.save(state, output_model_file)
        logger.info("Model weights saved in {}".format(output_model_file))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        config = kwargs.pop("config", None)
        state_dict = kwargs.pop("state_dict", None)
        kwargs.pop("cache_dir", None)
        kwargs.pop("from_tf", None)
        kwargs.pop("force_download", None)
        kwargs.pop("resume_download", None)
        kwargs.pop("proxies", None)
        output_loading_info = kwargs.pop("output_loading_info", False)
        kwargs.pop("local_files_only", None)
        kwargs.pop("use_cdn", None)
        kwargs.pop("error_on_mismatch", None)   # accepted and ignored, like the reference (:552)

        if config is None:
            config = cls.config_class.from_pretrained(pretrained_model_name_or_path, **kwargs)
            model_kwargs = {}
        elif isinstance(config, (str, os.PathLike)):
            config = cls.config_class.from_pretrained(config, **kwargs)
            model_kwargs = {}
        else:
            model_kwargs = kwargs

        if pretrained_model_name_or_path is not None and state_dict is None:
            if os.path.isdir(pretrained_model_name_or_path):
                archive_file = os.path.join(pretrained_model_name_or_path, WEIGHTS_NAME)
                if not os.path.isfile(archive_file):
                    raise EnvironmentError("Error no file named {} found in directory {}".format(
                        WEIGHTS_NAME, pretrained_model_name_or_path))
            elif os.path.isfile(pretrained_model_name_or_path):
                archive_file = pretrained_model_name_or_path
            else:
                raise EnvironmentError(
                    "Can't load weights for '{}': not a local directory or file (hub downloads need network "
                    "access, which this build does not have)".format(pretrained_model_name_or_path))
            try:
                state_dict = torch.load(archive_file, map_location="cpu")
            except Exception:
                raise OSError("Unable to load weights from pytorch checkpoint file.")

        model = cls(config, *model_args, **model_kwargs)
        missing_keys, unexpected_keys, error_msgs = [], [], []
        if state_dict is not None:
            # legacy LayerNorm names
            for key in list(state_dict.keys()):
                new_key = None
                if "gamma" in key:
                    new_key = key.replace("gamma", "weight")
                if "beta" in key:
                    new_key = key.replace("beta", "bias")
                if new_key:
                    state_dict[new_key] = state_dict.pop(key)
            has_prefix = any(s.startswith(cls.base_model_prefix + ".") for s in state_dict.keys())
            target = model
            prefix = ""
            if not hasattr(model, cls.base_model_prefix) and has_prefix:
                prefix = cls.base_model_prefix + "."
            if hasattr(model, cls.base_model_prefix) and not has_prefix:
                target = getattr(model, cls.base_model_prefix)
            partial = set(getattr(config, "partial_load", ()) or ())
            own = dict(target.state_dict())
            seen = set()
            with torch.no_grad():
                for name, dst in own.items():
                    key = prefix + name
                    if key not in state_dict:
                        missing_keys.append(name)
                        continue
                    seen.add(key)
                    src = state_dict[key]
                    if src.shape == dst.shape:
                        dst.copy_(src)
                    elif name in partial or key in partial:
                        # partial_load: copy the overlapping slice (src/model/mixins.py:511-528) — this is how
                        # facebook/bart-base (vocab 50265) loads into vocab 50320
                        if src.dim() != dst.dim() or any(s > d for s, d in zip(src.shape, dst.shape)):
                            error_msgs.append("size mismatch for {}: checkpoint {} vs model {}".format(key, tuple(src.shape), tuple(dst.shape)))
                        else:
                            dst[tuple(map(slice, src.size()))].copy_(src)
                    else:
                        error_msgs.append("size mismatch for {}: copying a param with shape {} from checkpoint, "
                                          "the shape in current model is {}.".format(key, tuple(src.shape), tuple(dst.shape)))
            unexpected_keys = [k for k in state_dict.keys() if k not in seen]
            if missing_keys:
                logger.info("Weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, missing_keys))
            if unexpected_keys:
                logger.info("Weights from pretrained model not used in {}: {}".format(model.__class__.__name__, unexpected_keys))
            if error_msgs:  # logged, never raised — same as the reference (:856-863)
                logger.warning("Error(s) in loading state_dict for {}:\n\t{}".format(model.__class__.__name__, "\n\t".join(error_msgs)))
        model.tie_weights()
        model.eval()
        if output_loading_info:
            return model, {"missing_keys": missing_keys, "unexpected_keys": unexpected_keys, "error_msgs": error_msgs}
        return model
