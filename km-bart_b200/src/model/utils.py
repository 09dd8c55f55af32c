"""sample_sentence — drop-in for the reference's src/model/utils.py:6-58 (no-cache top-k/top-p
sampler that also returns the summed log-probabilities of the sampled tokens)."""
import torch
from torch.nn import functional as F

from src.model.mixins import top_k_top_p_filtering


def sample_sentence(model, input_ids, image_features, attention_mask, tokenizer, top_k=50, top_p=1.0, max_length=20):
    batch_size = input_ids.shape[0]
    encoder = model.get_encoder()
    encoder_outputs = encoder(input_ids, image_features, attention_mask=attention_mask)

    unfinished_sents = input_ids.new(batch_size).fill_(1)
    sent_lengths = input_ids.new(batch_size).fill_(max_length)
    logprobs = []
    decoder_input_ids = input_ids.new(batch_size, 1).fill_(tokenizer.bos_token_id)

    cur_len = 1
    while cur_len < max_length:
        outputs = model(input_ids=None, decoder_input_ids=decoder_input_ids, image_features=image_features,
                        attention_mask=attention_mask, encoder_outputs=encoder_outputs, use_cache=False)
        next_token_logits = outputs[0][:, -1, :]
        next_token_logits = top_k_top_p_filtering(next_token_logits, top_k=top_k, top_p=top_p)
        next_token = torch.multinomial(F.softmax(next_token_logits, dim=-1), num_samples=1).squeeze(1)
        _scores = torch.gather(F.log_softmax(next_token_logits, dim=-1), -1, next_token.unsqueeze(-1))
        logprobs.append(_scores)

        tokens_to_add = next_token * unfinished_sents + tokenizer.pad_token_id * (1 - unfinished_sents)
        decoder_input_ids = torch.cat([decoder_input_ids, tokens_to_add.unsqueeze(-1)], dim=-1)
        cur_len = cur_len + 1
        eos_in_sents = tokens_to_add == tokenizer.eos_token_id
        newly_done = unfinished_sents.mul(eos_in_sents.long()).bool()
        sent_lengths.masked_fill_(newly_done, cur_len)
        unfinished_sents.mul_((~eos_in_sents).long())
        if unfinished_sents.max() == 0:
            break

    logprobs = torch.cat(logprobs, dim=1)
    for i in range(batch_size):
        logprobs[i, sent_lengths[i] - 1:] = 0
    return decoder_input_ids, logprobs.sum(dim=1).unsqueeze(1)
