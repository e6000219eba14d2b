"""Caller of the hot path: greedy / sampling decode loop with a k_v cache.

Mirrors the public surface of CleanTransformer/generation/generation_util.py (`GenerationMixin`
with `generate(input_ids, attention_mask, position_ids, segment_ids, generation_configs, steamers)`)
for beam_size == 1. The loop itself is host-side control flow (kept in Python); every
model call inside it runs on the sm_100a kernels. Loop semantics follow generation_util.py:57-119:
only the new tokens are fed once a cache exists, finished rows emit pad_id, the mask grows by
repeating its last column, and the loop ends when `step > max_gen_len + prompt_len` — i.e. it emits
max_gen_len + 2 tokens, exactly like the reference. Beam search (generation_util.py:207-290) is out
of scope for this tier.
"""
import torch


def _temperature(scores, temperature):
    """logits_processor.py:35-41 (TemperatureLogitsWrapper): the temperature is clamped at 1e-2."""
    return scores / max(temperature, 1e-2)


def _filter_top_k(scores, k, min_keep=1):
    """logits_processor.py:44-56 (TopKLogitsWrapper)."""
    k = min(int(max(k, min_keep, 1)), scores.size(-1))
    kth = torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(scores < kth, float("-inf"))


def _filter_top_p(scores, top_p, min_keep=1):
    """logits_processor.py:59-79 (TopPLogitsWrapper): top_p clamped to [0, 1]; the `min_keep` (>= 1) most probable
    tokens always survive — with top_p = 0 that is exactly the arg-max."""
    top_p = max(min(top_p, 1.0), 0)
    sorted_scores, sorted_idx = torch.sort(scores, descending=False)
    cum = sorted_scores.softmax(dim=-1).cumsum(dim=-1)
    remove = cum <= (1 - top_p)
    remove[..., -max(1, min_keep):] = False
    remove = remove.scatter(1, sorted_idx, remove)
    return scores.masked_fill(remove, float("-inf"))


class GenerationMixin:
    def generate(self, input_ids, attention_mask=None, position_ids=None, segment_ids=None,
                 generation_configs={}, steamers=None):
        cfg = generation_configs
        if cfg.get("beam_size", 1) != 1:
            raise NotImplementedError("beam search is outside the accelerated hot path (see DESIGN.md)")
        if cfg.get("no_repeat_ngram_size", 0) > 1:
            raise NotImplementedError("no_repeat_ngram processor is not part of the hot path")
        end_ids = cfg.get("end_ids", None)
        if isinstance(end_ids, int):
            end_ids = [end_ids]
        end_t = torch.tensor(list(end_ids), device=input_ids.device) if end_ids is not None else None
        return self._greedy_search(input_ids, attention_mask, position_ids, segment_ids, end_t,
                                   max_gen_len=cfg.get("max_gen_len", 100), pad_id=cfg.get("pad_id", 0),
                                   do_sample=cfg.get("do_sample", True), temperature=cfg.get("temperature", 1.0),
                                   top_k=cfg.get("top_k", 10), top_p=cfg.get("top_p", 0.8), steamers=steamers)

    @torch.no_grad()
    def _greedy_search(self, input_ids, attention_mask, position_ids, segment_ids, end_ids_tensor,
                       max_gen_len, pad_id, do_sample=False, temperature=1.0, top_k=0, top_p=1.0,
                       steamers=None):
        bsz, prompt_len = input_ids.shape
        limit = max_gen_len + prompt_len
        caches = [None] * self.config.n_layer
        alive = torch.ones(bsz, dtype=torch.long, device=input_ids.device)
        fed = 0  # number of tokens already inside the cache
        callbacks = [] if steamers is None else (steamers if isinstance(steamers, list) else [steamers])
        while True:
            kwargs = dict(attention_mask=attention_mask, k_v_pasts=caches)
            if position_ids is not None:
                kwargs["position_ids"] = position_ids[:, fed:]
            if segment_ids is not None:
                kwargs["segment_ids"] = segment_ids[:, fed:]
            outputs, caches = self(input_ids[:, fed:], **kwargs)
            scores = outputs[0][:, -1, :]
            if do_sample:
                scores = scores.float()
                if temperature != 1.0:
                    scores = _temperature(scores, temperature)
                if top_k > 0:
                    scores = _filter_top_k(scores, top_k)
                if top_p < 1.0:
                    scores = _filter_top_p(scores, top_p)
                nxt = torch.multinomial(torch.softmax(scores, dim=-1), num_samples=1).squeeze(1)
            else:
                nxt = torch.argmax(scores, dim=-1)
            nxt = nxt * alive + pad_id * (1 - alive)
            if end_ids_tensor is not None:
                hit = (nxt[None, :] == end_ids_tensor[:, None]).any(dim=0)
                alive = alive * (~hit).long()
            input_ids = torch.cat([input_ids, nxt[:, None]], dim=-1)
            if position_ids is not None:
                position_ids = torch.cat([position_ids, position_ids.max(dim=-1).values[:, None] + 1], dim=-1)
            if segment_ids is not None:
                segment_ids = torch.cat([segment_ids, segment_ids[:, -1:]], dim=-1)
            attention_mask = torch.cat([attention_mask, attention_mask[:, -1:]], dim=-1)
            stop = False
            for cb in callbacks:
                if callable(cb) and cb(input_ids.view(bsz, 1, -1)):
                    stop = True
            if stop:
                break
            fed = input_ids.shape[1] - 1
            if alive.max() == 0 or fed > limit:
                break
        return input_ids.view(bsz, 1, -1)
