"""Host-side mirror of CleanTransformer/models/modeling_gpt.py over the sm_100a kernels.

Same classes / parameter names as the reference (Conv1D weights stay [in, out]; the causal `bias`
buffer stays in the state_dict so HF-remapped checkpoints load strictly, examples/inference_gpt2.py
:16-41). Arithmetic per block:
    LayerNorm -> bf16 ; c_attn GEMM (MN-major B operand reads the [in,out] weight in place) ;
    attention kernel in its GPT variant: causal entries REPLACED by -1e4 (`w*b + -1e4*(1-b)`,
    modeling_gpt.py:88-89) then the additive finfo.min padding mask (:91-92, :176-179) ;
    c_proj GEMM + bias + residual epilogue ; MLP GEMM + gelu_new epilogue ; GEMM + residual.
"""
import math

import torch

from .. import functional as F
from .. import ops
from ..generation import GenerationMixin
from ..transformer import LayerNorm


class GPTConfig():
    """modeling_gpt.py:14-29."""

    def __init__(self, vocab_size=100, n_embd=100, n_positions=100, n_layer=3, n_head=2, n_ctx=2000,
                 embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1, layer_norm_epsilon=1e-5,
                 afn='gelu_new', **kwargs):
        self.vocab_size, self.n_embd, self.n_positions = vocab_size, n_embd, n_positions
        self.n_layer, self.n_head, self.n_ctx = n_layer, n_head, n_ctx
        self.embd_pdrop, self.attn_pdrop, self.resid_pdrop = embd_pdrop, attn_pdrop, resid_pdrop
        self.layer_norm_epsilon = layer_norm_epsilon
        self.afn = afn
        for k, v in kwargs.items():
            setattr(self, k, v)


class Conv1D(torch.nn.Module):
    """modeling_gpt.py:32-46: a Linear whose weight is stored [input_dim, out_dim]."""

    def __init__(self, out_dim, input_dim):
        super(Conv1D, self).__init__()
        w = torch.empty(input_dim, out_dim)
        torch.nn.init.normal_(w, std=0.02)
        self.weight = torch.nn.Parameter(w)
        self.bias = torch.nn.Parameter(torch.zeros(out_dim))

    def forward(self, x, act=ops.ACT_NONE, residual=None, out_dtype=None):
        return F.linear(x, self.weight, self.bias, act=act, residual=residual, out_dtype=out_dtype,
                        w_in_out=True)


class NewGELUActivation(torch.nn.Module):
    """modeling_gpt.py:112-122 (stand-alone; inside the MLP it is a GEMM epilogue)."""

    def forward(self, input):
        return _ActFn.apply(input, ops.ACT_GELU_TANH)


class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        ctx.save_for_backward(x)
        ctx.act = act
        return ops.act_fwd(x.detach(), act)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.act_bwd(g, x, ctx.act, out_dtype=g.dtype), None


class _Gelu(torch.nn.Module):
    def forward(self, x):
        return _ActFn.apply(x, ops.ACT_GELU_ERF)


class _Relu(torch.nn.Module):
    def forward(self, x):
        return _ActFn.apply(x, ops.ACT_RELU)


ACT2FN = {"gelu": _Gelu, "relu": _Relu, 'gelu_new': NewGELUActivation}
_ACT_ID = {"gelu": ops.ACT_GELU_ERF, "relu": ops.ACT_RELU, "gelu_new": ops.ACT_GELU_TANH}


class GptMask:
    """Per-key bias (log2 domain) + first valid key, built once per forward by GPTModel."""

    def __init__(self, kbias2, first_valid):
        self.kbias2, self.first_valid = kbias2, first_valid


def _mask_from_additive(attention_mask, bsz):
    """[b,1,1,t] additive mask as built at modeling_gpt.py:176-179 -> GptMask."""
    if attention_mask is None or isinstance(attention_mask, GptMask):
        return attention_mask
    m = attention_mask.reshape(attention_mask.shape[0], 1, attention_mask.shape[-1]).float()
    return GptMask((m.expand(bsz, 1, m.shape[-1]) * 1.4426950408889634).contiguous(), None)


class AttentionLayer(torch.nn.Module):
    """modeling_gpt.py:49-109."""

    def __init__(self, config, scale=False):
        super().__init__()
        self.config, self.scale = config, scale
        self.n_state, self.n_head, self.n_ctx = config.n_embd, config.n_head, config.n_ctx
        assert self.n_state % self.n_head == 0
        self.register_buffer("bias", torch.tril(torch.ones(self.n_ctx, self.n_ctx)).view(1, 1, self.n_ctx, self.n_ctx))
        self.c_attn = Conv1D(self.n_state * 3, self.n_state)
        self.c_proj = Conv1D(self.n_state, self.n_state)
        self.attn_dropout = torch.nn.Dropout(config.attn_pdrop)
        self.resid_dropout = torch.nn.Dropout(config.resid_pdrop)

    def forward(self, hidden_states, k_v_past=None, attention_mask=None, head_mask=None, residual=None):
        """`residual` (extension): when given, `residual + c_proj(ctx)` is produced by the GEMM
        epilogue and returned instead of the bare projection."""
        # modeling_gpt.py:93-96,107: dropout on the attention probabilities (inside the kernel) and on the projection
        a_drop = F.next_dropout(self.attn_dropout.p) if (self.training and self.attn_dropout.p > 0) else None
        r_on = self.training and self.resid_dropout.p > 0
        F.reject_head_mask(head_mask)
        bsz, q_len, _ = hidden_states.shape
        head_dim = self.n_state // self.n_head
        sm_scale = 1.0 / math.sqrt(head_dim) if self.scale else 1.0
        mask = _mask_from_additive(attention_mask, bsz)
        kb = mask.kbias2 if mask is not None else None
        fv = mask.first_valid if mask is not None else None
        qkv = self.c_attn(hidden_states)
        if isinstance(k_v_past, ops.StaticKV):
            # captured decode step (generation.py): append at the device-side position, attend over the device-side length
            q, k, v = F.split_packed(qkv, self.n_head, F.LAYOUT_GPT)
            ctx, _ = ops.attn_fwd(q, k_v_past.k, k_v_past.v, sm_scale, True, -1e4, kb, fv, need_lse=False,
                                  seq_len_dev=k_v_past.len_dev, kv_new=(k, v))  # the kernel appends k, v itself
            out = self.c_proj(ctx, residual=residual, out_dtype=None if residual is not None else torch.float32)
            return out, k_v_past
        if k_v_past is None and torch.is_grad_enabled() and qkv.requires_grad:
            ctx = F.PackedAttentionFn.apply(qkv, self.n_head, F.LAYOUT_GPT, sm_scale, True, -1e4, kb, fv, a_drop)
            _, k, v = F.split_packed(qkv.detach(), self.n_head, F.LAYOUT_GPT)
        else:
            q, k, v = F.split_packed(qkv, self.n_head, F.LAYOUT_GPT)
            if not torch.is_grad_enabled():
                # [b,h,t,d] cache (modeling_gpt.py:76-80) grown in place instead of torch.concat
                k = ops.kv_cache_append(None if k_v_past is None else k_v_past[0], k)
                v = ops.kv_cache_append(None if k_v_past is None else k_v_past[1], v)
            elif k_v_past is not None:
                k = torch.cat((k_v_past[0], k), dim=-2)
                v = torch.cat((k_v_past[1], v), dim=-2)
            ctx = F.attention_cached(q, k, v, sm_scale, True, -1e4, kb, fv, a_drop)
        if r_on:  # residual + dropout(c_proj(ctx)) (the caller's add, TransformerBlock.forward) as one kernel
            out = F.dropout(self.c_proj(ctx, out_dtype=torch.float32), self.resid_dropout.p, True, residual=residual)
        else:
            out = self.c_proj(ctx, residual=residual, out_dtype=None if residual is not None else torch.float32)
        return out, (k, v)


class TransformerBlock(torch.nn.Module):
    """modeling_gpt.py:125-153 (version 'gpt' = post-LN GPT-1, otherwise pre-LN GPT-2/3)."""

    def __init__(self, config, scale=False, version='gpt'):
        super(TransformerBlock, self).__init__()
        n_embd = config.n_embd
        self.version = version
        self.afn = config.afn
        self.attn = AttentionLayer(config, scale)
        self.norm1 = LayerNorm(n_embd, eps=config.layer_norm_epsilon)
        self.mlp = torch.nn.Sequential(
            Conv1D(4 * n_embd, n_embd),
            ACT2FN[config.afn](),
            Conv1D(n_embd, 4 * n_embd),
            torch.nn.Dropout()
        )
        self.norm2 = LayerNorm(n_embd, eps=config.layer_norm_epsilon)

    def _mlp(self, x, residual):
        h = self.mlp[0](x, act=_ACT_ID[self.afn])
        if self.training and self.mlp[3].p > 0:  # modeling_gpt.py:136: torch.nn.Dropout() (p = 0.5) closes the MLP
            return F.dropout(self.mlp[2](h, out_dtype=torch.float32), self.mlp[3].p, True, residual=residual)
        return self.mlp[2](h, residual=residual)

    def forward(self, x, attn_output=None, attention_mask=None, head_mask=None, k_v_past=None):
        cd = F.compute_dtype()
        x = x if x.dtype == torch.float32 else x.float()
        if self.version == 'gpt':
            if attn_output is None:
                s1, k_v_past = self.attn(x, attention_mask=attention_mask, head_mask=head_mask,
                                         k_v_past=k_v_past, residual=x)
            else:
                s1 = x + attn_output
            n1, n1_low = self.norm1(s1, out_dtype=torch.float32, out2_dtype=cd)
            output = self.norm2(self._mlp(n1_low, n1))
        else:
            if attn_output is None:
                x, k_v_past = self.attn(self.norm1(x, out_dtype=cd), attention_mask=attention_mask,
                                        head_mask=head_mask, k_v_past=k_v_past, residual=x)
            else:
                x = x + attn_output
            output = self._mlp(self.norm2(x, out_dtype=cd), x)
        return output, k_v_past


class GPTModel(torch.nn.Module):
    """modeling_gpt.py:156-195."""

    def __init__(self, config, version='gpt'):
        super(GPTModel, self).__init__()
        self.version = version
        self.config = config
        self.tokens_embed = torch.nn.Embedding(config.vocab_size, config.n_embd)
        self.position_embed = torch.nn.Embedding(config.n_positions, config.n_embd)
        self.drop = torch.nn.Dropout(config.embd_pdrop)
        self.blocks = torch.nn.ModuleList([TransformerBlock(config, scale=True, version=version)
                                           for _ in range(config.n_layer)])
        if version != 'gpt':
            self.ln_f = LayerNorm(config.n_embd, eps=config.layer_norm_epsilon)

    def forward(self, input_ids, attention_mask=None, position_ids=None, segment_ids=None, k_v_pasts=None):
        q_len = input_ids.shape[1]
        if position_ids is None:  # modeling_gpt.py:171-174 (index bookkeeping, not arithmetic)
            position_ids = attention_mask.long().cumsum(-1) - 1
            position_ids.masked_fill_(attention_mask == 0, 1)
            position_ids = position_ids[:, -q_len:]
        mask = None
        if isinstance(attention_mask, GptMask):
            mask = attention_mask  # prepared once by the caller (captured decode step: position_ids are given too)
        elif attention_mask is not None:
            kb, fv = ops.attn_mask_prep(attention_mask, self.config.n_head, ops.MASK_GPT)
            mask = GptMask(kb, fv)
        if k_v_pasts is None:
            k_v_pasts = [None] * len(self.blocks)
        ids, tables = [input_ids, position_ids], [self.tokens_embed.weight, self.position_embed.weight]
        if segment_ids is not None:
            ids.append(segment_ids.view(-1, segment_ids.size(-1)))
            tables.append(self.tokens_embed.weight)
        hidden_states = F.dropout(F.embedding_sum(ids, tables), self.drop.p, self.training and self.drop.p > 0)
        for i, block in enumerate(self.blocks):
            hidden_states, k_v_pasts[i] = block(hidden_states, attention_mask=mask, k_v_past=k_v_pasts[i])
        if self.version == 'gpt':
            return hidden_states, k_v_pasts
        return self.ln_f(hidden_states), k_v_pasts


class GPTLMHeadModel(torch.nn.Module, GenerationMixin):
    """modeling_gpt.py:198-214."""

    def __init__(self, config, version='gpt'):
        super(GPTLMHeadModel, self).__init__()
        self.config = config
        self.version = version
        self.gpt = GPTModel(config, version=version)
        self.lm_head = torch.nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self._tie_weights()

    def _tie_weights(self):
        self.lm_head.weight = self.gpt.tokens_embed.weight
        self.lm_head.weight._ct_expected_writes = 2  # lm_head wgrad + embedding scatter (see ddp.py)
        self.lm_head.weight._ct_sparse_second_write = True  # ... the second one being the token scatter

    _ct_graph_decode = True  # generation.py may replay the q_len = 1 step from a CUDA graph

    def _decode_static_mask(self, full_mask):
        """Per-key bias over the whole cache capacity (generated positions are valid keys), built once per generation."""
        kb, fv = ops.attn_mask_prep(full_mask, self.config.n_head, ops.MASK_GPT)
        return GptMask(kb, fv)

    def _decode_needs_positions(self):
        return True

    def _decode_graph_ok(self):
        return (self.config.n_embd // self.config.n_head) in (32, 64, 128)  # csrc/attention.cu: attn_decode_kernel<D>

    def forward(self, input_ids, attention_mask=None, segment_ids=None, position_ids=None, k_v_pasts=None):
        hidden_states, k_v_pasts = self.gpt(input_ids, attention_mask, position_ids, segment_ids, k_v_pasts)
        # fp32 logits: greedy token ids must be bit-exact vs the reference (BASELINE.json north_star)
        lm_logits = F.linear(hidden_states, self.lm_head.weight, out_dtype=torch.float32)
        return (lm_logits, hidden_states), k_v_pasts
