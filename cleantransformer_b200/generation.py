"""Caller of the hot path: greedy / sampling decode loop with a k_v cache.

Mirrors the public surface of CleanTransformer/generation/generation_util.py (`GenerationMixin`
with `generate(input_ids, attention_mask, position_ids, segment_ids, generation_configs, steamers)`)
for beam_size == 1. The loop itself is host-side control flow (kept in Python); every
model call inside it runs on the sm_100a kernels. Loop semantics follow generation_util.py:57-119:
only the new tokens are fed once a cache exists, finished rows emit pad_id, the mask grows by
repeating its last column, and the loop ends when `step > max_gen_len + prompt_len` — i.e. it emits
max_gen_len + 2 tokens, exactly like the reference. The no-repeat-ngram processor (logits_processor.py:11-32; what
examples/inference_bloom.py:93 configures) is one vectorised mask per step, and beam search
(generation_util.py:121-290; examples/inference_gpt2.py:64) is reproduced with its bookkeeping on the host — one small
read-back per step instead of the reference's `.item()` per candidate; both run the host loop (their next step
depends on the generated history), every model call inside it on the sm_100a kernels.

Decoding on a CUDA model in eval mode — greedy (`do_sample=False`) or sampling (the reference's default: temperature,
top-k, top-p, multinomial) — replays every q_len = 1 step from ONE captured CUDA graph (BASELINE.json
configs[3]): `_graphed_greedy` below. The cache length, the write position and the
alive flags live in device memory and are advanced by ct_greedy_step at the end of the step, so the graph's arguments
never change; the host only replays and (when end ids are given) polls a done flag every few steps.
"""
import os

import torch

from . import ops

POLL_EVERY = 16  # graphed greedy decode: steps between two reads of the device-side "every row finished" flag


def _on_device(t):
    return t.is_cuda


def _capture(step):
    """Capture one call of `step` (it is NOT executed) and return (replay, kernel launches per replay)."""
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    before = ops.LAUNCHES[0]
    with torch.cuda.graph(graph):
        step()
    n = ops.LAUNCHES[0] - before
    ops.LAUNCHES[0] = before
    return graph.replay, n


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


def _ban_repeated_ngrams(input_ids, scores, n):
    """logits_processor.py:11-32 (NoRepeatNGramLogitsProcessor): a token that would complete an n-gram already present
    in the row (padding included, like the reference) gets -inf. No host loop: every window of the row is compared
    with its last n-1 tokens at once."""
    bsz, cur = input_ids.shape
    if cur < n:
        return scores
    windows = input_ids.unfold(1, n, 1)                              # [bsz, cur-n+1, n]
    tail = input_ids[:, cur - n + 1:]                                # the n-1 tokens the next one would follow
    hit = (windows[:, :, :-1] == tail[:, None, :]).all(dim=-1)       # [bsz, cur-n+1]
    banned = torch.zeros(scores.shape, dtype=torch.int32, device=scores.device)
    banned.scatter_add_(1, windows[:, :, -1], hit.to(torch.int32))
    return scores.masked_fill(banned > 0, float("-inf"))


DECODE_PLAN_CACHE = [os.environ.get("CT_DECODE_CACHE", "1") != "0"]  # keep the captured step between generate() calls


class _DecodePlan:
    """A captured decode step and everything it refers to by address."""
    __slots__ = ("key", "fingerprint", "ids_out", "alive", "cur_ids", "pos_ids", "state", "end_ids", "mask_obj",
                 "static_kv", "replay", "launches")


def _fingerprint(model):
    """The graph reads the parameters through cached low-precision copies: any update (version) or move (address) of a
    parameter invalidates a cached plan."""
    return tuple((p._version, p.data_ptr()) for p in model.parameters())


def _make_sampler(temperature, top_k, top_p):
    """generation_util.py:74-84 for do_sample=True: logits wrappers (temperature, top-k, top-p) then one multinomial
    draw per row. Plain torch ops: the same closure serves the host loop and — they are all capturable — the captured
    decode step."""
    def sample(scores):
        scores = scores.float()
        if temperature != 1.0:
            scores = _temperature(scores, temperature)
        if top_k > 0:
            scores = _filter_top_k(scores, top_k)
        if top_p < 1.0:
            scores = _filter_top_p(scores, top_p)
        return torch.multinomial(torch.softmax(scores, dim=-1), num_samples=1).squeeze(1)
    sample.key = (float(temperature), int(top_k), float(top_p))
    return sample


class GenerationMixin:
    def generate(self, input_ids, attention_mask=None, position_ids=None, segment_ids=None,
                 generation_configs={}, steamers=None):
        cfg = generation_configs
        end_ids = cfg.get("end_ids", None)
        if isinstance(end_ids, int):
            end_ids = [end_ids]
        end_t = torch.tensor(list(end_ids), device=input_ids.device) if end_ids is not None else None
        common = dict(max_gen_len=cfg.get("max_gen_len", 100), pad_id=cfg.get("pad_id", 0),
                      do_sample=cfg.get("do_sample", True), temperature=cfg.get("temperature", 1.0),
                      top_k=cfg.get("top_k", 10), top_p=cfg.get("top_p", 0.8), steamers=steamers,
                      no_repeat_ngram_size=cfg.get("no_repeat_ngram_size", 0))
        if cfg.get("beam_size", 1) == 1:
            return self._greedy_search(input_ids, attention_mask, position_ids, segment_ids, end_t, **common)
        return self._beam_search(input_ids, attention_mask, position_ids, segment_ids, end_t,
                                 beam_size=cfg["beam_size"], early_stop=cfg.get("early_stop", True), **common)

    @torch.no_grad()
    def _greedy_search(self, input_ids, attention_mask, position_ids, segment_ids, end_ids_tensor,
                       max_gen_len, pad_id, do_sample=False, temperature=1.0, top_k=0, top_p=1.0,
                       steamers=None, no_repeat_ngram_size=0):
        bsz, prompt_len = input_ids.shape
        limit = max_gen_len + prompt_len
        sampler = _make_sampler(temperature, top_k, top_p) if do_sample else None
        if (steamers is None and _on_device(input_ids) and position_ids is None and segment_ids is None
                and attention_mask is not None and getattr(self, "_ct_graph_decode", False) and self._decode_graph_ok()
                and os.environ.get("CT_DECODE_GRAPH", "1") != "0" and max_gen_len >= 1
                and not self.training  # (train mode may have dropout active: its mask counter is host state)
                and no_repeat_ngram_size <= 1  # (the banned set depends on the generated history: host loop)
                and bool((attention_mask[:, -1] != 0).all())):
            return self._graphed_greedy(input_ids, attention_mask, end_ids_tensor, max_gen_len, pad_id, sampler)
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
            if no_repeat_ngram_size > 1:
                scores = _ban_repeated_ngrams(input_ids, scores, no_repeat_ngram_size)
            if do_sample:
                nxt = sampler(scores)
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

    @torch.no_grad()
    def _beam_search(self, input_ids, attention_mask, position_ids, segment_ids, end_ids_tensor, max_gen_len, pad_id,
                     beam_size, early_stop=True, do_sample=False, temperature=1.0, top_k=0, top_p=1.0, steamers=None,
                     no_repeat_ngram_size=0, length_penalty=1.0):
        """generation_util.py:121-290, behaviour for behaviour: every row is expanded to `beam_size` beams (only beam 0
        is live at the first step), 2 x beam_size continuations are drawn per row from the joint beam x vocabulary
        scores (arg-top-k, or the logits wrappers + multinomial without replacement when sampling: the wrappers then act
        on the JOINT distribution and the beam scores enter multiplied by the temperature), the first beam_size of
        them are examined: an end id closes a candidate (score = log-probability / length ** length_penalty; the
        beam_size best are kept), anything else fills the next live slot — slots that stay empty keep beam 0 / token 0
        / score 0, as in the reference. A row is done after beam_size candidates (early_stop) or once its worst kept
        candidate beats the best score still reachable; done rows emit pad_id. The loop runs until
        `step > max_gen_len + prompt_len` (it does not stop early) and returns the LIVE beams
        [bsz, beam_size, prompt + max_gen_len + 2]. Like the reference it needs end ids."""
        if end_ids_tensor is None:
            raise TypeError("beam search needs generation_configs['end_ids'] (the reference tests every candidate "
                            "against them, generation_util.py:140)")
        bsz, prompt_len = input_ids.shape
        limit = max_gen_len + prompt_len
        dev = input_ids.device
        B = beam_size

        def spread(t):
            return None if t is None else t.repeat_interleave(B, dim=0)

        input_ids, position_ids, attention_mask, segment_ids = (spread(input_ids), spread(position_ids),
                                                                spread(attention_mask), spread(segment_ids))
        beam_scores = torch.zeros(bsz, B, device=dev)
        beam_scores[:, 1:] = -1e9
        ends = set(int(e) for e in end_ids_tensor.reshape(-1).tolist())
        rows = [dict(done=False, worst=torch.tensor(1e9, device="cpu"), kept=[]) for _ in range(bsz)]   # kept: candidate scores
        caches = [None] * self.config.n_layer
        callbacks = [] if steamers is None else (steamers if isinstance(steamers, list) else [steamers])
        fed = 0
        while True:
            kwargs = dict(attention_mask=attention_mask, k_v_pasts=caches)
            if position_ids is not None:
                kwargs["position_ids"] = position_ids[:, fed:]
            if segment_ids is not None:
                kwargs["segment_ids"] = segment_ids[:, fed:]
            outputs, caches = self(input_ids[:, fed:], **kwargs)
            scores = outputs[0][:, -1, :].float()
            if no_repeat_ngram_size > 1:
                scores = _ban_repeated_ngrams(input_ids, scores.view(bsz * B, -1), no_repeat_ngram_size)
            # -- the 2B best continuations of every row (generation_util.py:183-205)
            V = scores.shape[-1]
            joint = torch.log_softmax(scores, dim=-1) + \
                beam_scores.view(-1, 1) * (temperature if do_sample else 1.0)
            joint = joint.view(bsz, B * V)
            if do_sample:
                if temperature != 1.0:
                    joint = _temperature(joint, temperature)
                if top_k > 0:
                    joint = _filter_top_k(joint, top_k)
                if top_p < 1.0:
                    joint = _filter_top_p(joint, top_p)
                drawn = torch.multinomial(torch.softmax(joint, dim=-1), num_samples=2 * B)
                cand_scores, order = torch.sort(torch.gather(joint, -1, drawn), descending=True, dim=1)
                drawn = torch.gather(drawn, -1, order)
            else:
                cand_scores, drawn = joint.topk(2 * B, dim=1, largest=True, sorted=True)
            from_beam = torch.div(drawn, V, rounding_mode="floor")
            tokens = drawn % V
            # -- bookkeeping on the host (generation_util.py:121-181): one read-back per step
            h_beam, h_tok, h_score = from_beam.cpu(), tokens.cpu(), cand_scores.float().cpu()
            cur_len = input_ids.shape[-1]
            new_beam = torch.zeros(bsz, B, dtype=from_beam.dtype, device="cpu")   # (explicit: run.py makes the GPU the default)
            new_tok = torch.zeros(bsz, B, dtype=tokens.dtype, device="cpu")
            new_score = torch.zeros(bsz, B, dtype=cand_scores.dtype, device="cpu")
            for r, row in enumerate(rows):
                if row["done"]:
                    new_tok[r, :] = pad_id
                    continue
                live = 0
                for c in range(B):
                    if int(h_tok[r, c]) in ends:
                        sc = h_score[r, c] / (cur_len ** length_penalty)
                        row["kept"].append(sc)
                        if len(row["kept"]) > B:
                            ranked = sorted((float(v), i) for i, v in enumerate(row["kept"]))
                            del row["kept"][ranked[0][1]]
                            row["worst"] = torch.tensor(ranked[1][0], device="cpu")
                        else:
                            row["worst"] = torch.minimum(sc, row["worst"])
                    else:
                        new_beam[r, live], new_tok[r, live], new_score[r, live] = h_beam[r, c], h_tok[r, c], h_score[r, c]
                        live += 1
                    if live >= B:
                        break
                if len(row["kept"]) >= B:
                    if early_stop:
                        row["done"] = True
                        continue
                    reachable = float(h_score[r].max()) / ((cur_len + 1) ** length_penalty)
                    if float(row["worst"]) > reachable:
                        row["done"] = True
            from_beam, tokens, beam_scores = new_beam.to(dev), new_tok.to(dev), new_score.to(dev)

            # -- the chosen beams become the next inputs; their caches follow them
            def follow(value, how):
                if value is None:
                    return None
                value = value.view(bsz, B, -1)
                value = value.gather(1, from_beam[:, :, None].expand_as(value)).view(bsz * B, -1)
                if how == "token":
                    return torch.cat([value, tokens.view(-1)[:, None]], dim=-1)
                if how == "position":
                    return torch.cat([value, value[:, -1:] + 1], dim=-1)
                return torch.cat([value, value[:, -1:]], dim=-1)

            input_ids = follow(input_ids, "token")
            position_ids = follow(position_ids, "position")
            attention_mask = follow(attention_mask, "same")
            segment_ids = follow(segment_ids, "same")
            flat = (from_beam + torch.arange(bsz, device=dev)[:, None] * B).view(-1)
            caches = [tuple(t.index_select(0, flat) for t in layer) for layer in caches]
            stop = False
            for cb in callbacks:
                if callable(cb) and cb(input_ids.view(bsz, B, -1)):
                    stop = True
            if stop:
                break
            fed = input_ids.shape[1] - 1
            if fed > limit:
                break
        return input_ids.view(bsz, B, -1)

    @torch.no_grad()
    def _graphed_greedy(self, input_ids, attention_mask, end_ids_tensor, max_gen_len, pad_id, sampler=None):
        """Same token ids as the loop above for do_sample=False (generation_util.py:57-119), with the q_len = 1 steps
        replayed from a CUDA graph. With a `sampler` (do_sample=True) the draw — torch's own processors and multinomial,
        captured with the step; torch advances the generator's Philox offset per replay — replaces the argmax; the
        bookkeeping stays in ct_greedy_step. The prompt's last mask column is 1 for every row (checked by the caller), so
        every generated position is a valid key and its GPT position id is the previous one + 1.

        Step k >= 1 of the reference feeds token P+k-1 and emits token P+k; it stops after the step that leaves
        `fed > max_gen_len + P`, i.e. after max_gen_len + 2 emitted tokens, or right after the step in which the last
        row hit an end id. Cache rows needed: P + max_gen_len + 1.

        The captured step and every buffer it addresses (KV caches, mask bias, counters, output ids) form a `_DecodePlan`
        kept on the model: the next generate() with the same shapes / options and unchanged parameters re-initialises the
        buffers in place and replays — no second capture (10 - 200 ms of host time per generation, r03f)."""
        dev = input_ids.device
        bsz, P = input_ids.shape
        n_emit = max_gen_len + 2
        cap = P + max_gen_len + 1
        n_end = 0 if end_ids_tensor is None else int(end_ids_tensor.numel())
        key = (bsz, P, int(max_gen_len), int(pad_id), n_end, getattr(sampler, "key", None), str(dev))
        plan = getattr(self, "_ct_decode_plan", None)
        reuse = (plan is not None and plan.key == key and DECODE_PLAN_CACHE[0]
                 and plan.fingerprint == _fingerprint(self))
        self._ct_decode_plan_reused = bool(reuse)
        full_mask = torch.cat([attention_mask, attention_mask[:, -1:].expand(bsz, cap - P)], dim=-1).contiguous()
        state0 = torch.tensor([P, P, bsz, -1, 0], dtype=torch.int32)
        if not reuse:
            self._ct_decode_plan = plan = None  # (frees the previous plan's caches and graph before allocating new ones)
            plan = _DecodePlan()
            plan.key = key
            plan.ids_out = torch.empty(bsz, P + n_emit, dtype=torch.long, device=dev)
            plan.alive = torch.ones(bsz, dtype=torch.long, device=dev)
            plan.cur_ids = torch.empty(bsz, dtype=torch.long, device=dev)
            # position of the last prompt token; ct_greedy_step adds 1 per emitted token
            plan.pos_ids = torch.empty(bsz, dtype=torch.long, device=dev) if self._decode_needs_positions() else None
            # state: cache length seen by the next step, write column, alive rows, done_at, block counter
            plan.state = torch.empty(5, dtype=torch.int32, device=dev)
            plan.end_ids = None if n_end == 0 else torch.empty(n_end, dtype=torch.long, device=dev)
            plan.mask_obj = self._decode_static_mask(full_mask)
        else:
            fresh = self._decode_static_mask(full_mask)
            for name in ("kbias2", "first_valid"):
                dst, src = getattr(plan.mask_obj, name, None), getattr(fresh, name, None)
                if dst is not None:
                    dst.copy_(src)
            plan.alive.fill_(1)
        plan.ids_out[:, :P] = input_ids
        plan.state.copy_(state0)
        if plan.pos_ids is not None:
            plan.pos_ids.copy_(attention_mask.long().cumsum(-1)[:, -1] - 1)
        if plan.end_ids is not None:
            plan.end_ids.copy_(end_ids_tensor.to(device=dev, dtype=torch.long).reshape(-1))
        ids_out, alive, cur_ids, pos_ids, state, end_ids = (plan.ids_out, plan.alive, plan.cur_ids, plan.pos_ids,
                                                            plan.state, plan.end_ids)

        def pick(logits):
            drawn = None if sampler is None else sampler(logits).contiguous()
            ops.greedy_step(logits, alive, end_ids, pad_id, ids_out, cur_ids, pos_ids, state, sampled=drawn)

        trace = getattr(self, "_ct_decode_trace", None)  # tools/decode_timing.py: a list that receives phase timings
        marks = []

        def mark(name):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        # prefill: the un-graphed full-sequence path; the caches are allocated once for the whole generation, or — with a
        # cached plan — written into the plan's buffers (ops.KV_PREALLOC hands them to kv_cache_append in layer order)
        old_cap = ops.KV_CACHE_MIN_CAP[0]
        ops.KV_CACHE_MIN_CAP[0] = cap
        ops.KV_PREALLOC.clear()
        if reuse:
            for kv in plan.static_kv:
                ops.KV_PREALLOC.extend((kv.k, kv.v))
        try:
            outputs, caches = self(input_ids, attention_mask=attention_mask, k_v_pasts=[None] * self.config.n_layer)
        finally:
            ops.KV_CACHE_MIN_CAP[0] = old_cap
            ops.KV_PREALLOC.clear()
        if reuse and not all(k._ct_cache_base is kv.k and v._ct_cache_base is kv.v
                             for (k, v), kv in zip(caches, plan.static_kv)):
            raise RuntimeError("generate(): the prefill did not land in the cached decode plan's KV buffers")
        pick(outputs[0][:, -1, :])
        if not reuse:
            plan.static_kv = [ops.StaticKV(k._ct_cache_base, v._ct_cache_base, state) for k, v in caches]
        static_kv, mask_obj = plan.static_kv, plan.mask_obj

        def step():
            kw = dict(attention_mask=mask_obj, k_v_pasts=static_kv)
            if pos_ids is not None:
                kw["position_ids"] = pos_ids.view(bsz, 1)
            out, _ = self(cur_ids.view(bsz, 1), **kw)
            pick(out[0][:, -1, :])

        def finished():
            return end_ids is not None and int(state[3]) >= 0

        n_steps = n_emit - 1  # q_len = 1 steps still to run
        mark("prefill")
        # programmatic dependent launch for the ~170 few-microsecond kernels of a decode step (csrc/ct_common.cuh): +9 %;
        # the option is left alone when the user forced it on or off (CT_PDL = 1 / 2)
        pdl_prev = ops.set_option("PDL", 1) if ops.get_option("PDL") == 0 else None
        try:
            if not reuse:
                if n_steps > 0 and not finished():
                    step()  # first decode step outside the capture: lazy kernel attributes, and it is a real step
                    n_steps -= 1
                mark("first_step")
                plan.replay, plan.launches = None, 0
                if n_steps > 0 and not finished():
                    plan.replay, plan.launches = _capture(step)
                    plan.fingerprint = _fingerprint(self)
                    if DECODE_PLAN_CACHE[0]:
                        self._ct_decode_plan = plan  # (one plan per model: its caches and graph stay alive)
                mark("capture")
            self._ct_decode_graph_launches = plan.launches
            if plan.replay is not None and n_steps > 0 and not finished():
                import time
                t_issue = time.perf_counter()
                done = 0
                while done < n_steps:
                    burst = min(POLL_EVERY, n_steps - done) if end_ids is not None else n_steps - done
                    for _ in range(burst):
                        plan.replay()
                    done += burst
                    ops.LAUNCHES[0] += burst * plan.launches
                    if finished():
                        break
                mark("replays")
                if trace is not None:  # host time to ISSUE the replays: far below the device time = the host is not the limit
                    self._ct_replay_issue_ms = (time.perf_counter() - t_issue) * 1e3
        finally:
            if pdl_prev is not None:
                ops.set_option("PDL", pdl_prev)
        st = state.tolist()
        if trace is not None:
            torch.cuda.synchronize()
            trace.append({b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])})
            trace[-1]["replay_issue_host_ms"] = getattr(self, "_ct_replay_issue_ms", None)
            trace[-1]["plan_reused"] = bool(reuse)
        n_out = st[3] if (end_ids is not None and st[3] >= 0) else min(st[1], P + n_emit)
        return ids_out[:, :n_out].clone().reshape(bsz, 1, -1)  # (a copy: the buffer belongs to the plan)
