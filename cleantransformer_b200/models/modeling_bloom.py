"""Host-side mirror of CleanTransformer/models/modeling_bloom.py over the sm_100a kernels.

Class names, constructor arguments, sub-module / parameter names and forward signatures follow the
reference (so `load_state_dict(strict=True)` with reference / HF-remapped checkpoints works,
examples/inference_bloom.py:16-47). The per-layer arithmetic is:
    LayerNorm (ct_layernorm)  -> bf16
    fused QKV GEMM + bias (tcgen05)                       modeling_bloom.py:79
    flash-style attention with ALiBi + causal + key padding, reading the per-head interleaved
    [b,s,h,3,d] projection in place                        modeling_bloom.py:81-116
    dense GEMM + bias + residual epilogue -> f32 stream    modeling_bloom.py:121-122
    LayerNorm -> bf16
    h->4h GEMM + bias + tanh-GELU epilogue (keeps the pre-activation for backward)  :256
    4h->h GEMM + bias + residual epilogue                  modeling_bloom.py:267-269
The tril/alibi tensors of the reference (built on the CPU every step, :181-182,193) are replaced by
one tiny kernel that turns attention_mask into a per-key bias (ops.attn_mask_prep).
"""
import math

import torch

from .. import functional as F
from .. import ops
from ..generation import GenerationMixin
from ..transformer import LayerNorm

LOG2E = 1.4426950408889634


class BloomConfig():
    """modeling_bloom.py:17-54 (same keyword set, incl. the n_embed synonym)."""

    def __init__(self, vocab_size=250880, hidden_size=64, n_layer=2, num_attention_heads=8,
                 layer_norm_epsilon=1e-5, initializer_range=0.02, use_cache=True, bos_token_id=1,
                 eos_token_id=2, apply_residual_connection_post_layernorm=False, hidden_dropout=0.0,
                 attention_dropout=0.0, pretraining_tp=1, slow_but_exact=False, **kwargs):
        n_embed = kwargs.pop("n_embed", None)
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size if n_embed is None else n_embed
        self.n_layer = self.num_hidden_layers = n_layer
        self.n_head = self.num_attention_heads = num_attention_heads
        self.layer_norm_epsilon = layer_norm_epsilon
        self.initializer_range = initializer_range
        self.use_cache = use_cache
        self.pretraining_tp = pretraining_tp
        self.apply_residual_connection_post_layernorm = apply_residual_connection_post_layernorm
        self.hidden_dropout = hidden_dropout
        self.attention_dropout = attention_dropout
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.slow_but_exact = slow_but_exact


_SLOPES = {}


def alibi_slopes(num_heads, device=None):
    """Per-head ALiBi slopes, modeling_bloom.py:313-326 — the reference's own sequence of fp32 tensor operations
    (torch.pow of an fp32 base by int32 powers, on the mask's device), so the values are the ones the reference
    multiplies with, not a higher-precision restatement that differs in the last bit for 12, 16, 20 … heads.
    Cached per (heads, device)."""
    key = (num_heads, str(device))
    hit = _SLOPES.get(key)
    if hit is not None:
        return hit
    closest = 2 ** math.floor(math.log2(num_heads))
    base = torch.tensor(2 ** (-(2 ** -(math.log2(closest) - 3))), device=device, dtype=torch.float32)
    slopes = torch.pow(base, torch.arange(1, 1 + closest, device=device, dtype=torch.int32))
    if closest != num_heads:
        extra = torch.tensor(2 ** (-(2 ** -(math.log2(2 * closest) - 3))), device=device, dtype=torch.float32)
        n_rem = min(closest, num_heads - closest)
        slopes = torch.cat([slopes, torch.pow(extra, torch.arange(1, 1 + 2 * n_rem, 2, device=device, dtype=torch.int32))], dim=0)
    _SLOPES[key] = slopes
    return slopes


def build_alibi_tensor(attention_mask, num_heads, dtype):
    """modeling_bloom.py:309-331 — kept for API parity (the attention kernel consumes the
    equivalent per-key bias from ops.attn_mask_prep instead)."""
    batch_size, seq_length = attention_mask.shape
    slopes = alibi_slopes(num_heads, attention_mask.device)
    pos = ((attention_mask.cumsum(dim=-1) - 1) * attention_mask)[:, None, :]
    return (slopes[..., None] * pos).reshape(batch_size * num_heads, 1, seq_length).to(dtype)


class AttnBias:
    """What BloomModel hands to its blocks in place of the reference's (alibi, bool-mask) pair."""

    def __init__(self, kbias2, first_valid, causal):
        self.kbias2, self.first_valid, self.causal = kbias2, first_valid, causal

    @staticmethod
    def from_mask(attention_mask, n_head, q_len):
        kb, fv = ops.attn_mask_prep(attention_mask, n_head, ops.MASK_BLOOM,
                                    alibi_slopes(n_head, attention_mask.device))
        return AttnBias(kb, fv, q_len > 1)

    @staticmethod
    def from_reference_args(alibi, attention_mask, bsz, n_head):
        """Accept the reference's tensors: alibi [b*h,1,k] and bool mask [b,1,q,k] (True = masked,
        causal part included as built by BloomModel._attn_mask, modeling_bloom.py:176-185)."""
        k_len = alibi.shape[-1]
        kb = alibi.reshape(bsz, n_head, k_len).float() * LOG2E
        q_len = 1
        if attention_mask is not None:
            q_len = attention_mask.shape[2]
            key_masked = attention_mask[:, 0, -1, :]  # the last query row sees every non-future key
            kb = kb + torch.where(key_masked, float("-inf"), 0.0)[:, None, :]
        return AttnBias(kb.contiguous(), None, q_len > 1)


class BloomAttentionLayer(torch.nn.Module):
    """modeling_bloom.py:57-124."""

    def __init__(self, config):
        super().__init__()
        self.pretraining_tp = config.pretraining_tp
        self.slow_but_exact = config.slow_but_exact
        self.hidden_size = config.hidden_size
        self.num_heads = config.n_head
        self.head_dim = self.hidden_size // self.num_heads
        self.hidden_dropout = config.hidden_dropout
        self.inv_norm_factor = 1.0 / math.sqrt(self.head_dim)
        self.beta = 1.0
        self.query_key_value = torch.nn.Linear(self.hidden_size, 3 * self.hidden_size, bias=True)
        self.dense = torch.nn.Linear(self.hidden_size, self.hidden_size)
        self.attention_dropout = torch.nn.Dropout(config.attention_dropout)

    def forward(self, hidden_states, residual, alibi, k_v_past=None, attention_mask=None, head_mask=None):
        if self.pretraining_tp > 1 and self.slow_but_exact:
            raise Exception("pretraining_tp and slow_but_exact not supported yet")
        # modeling_bloom.py:111 (probabilities, inside the kernel) and :121-123 (residual + dropout(dense))
        a_drop = F.next_dropout(self.attention_dropout.p) if (self.training and self.attention_dropout.p > 0) else None
        h_on = self.training and self.hidden_dropout > 0
        F.reject_head_mask(head_mask)
        bsz, q_len, _ = hidden_states.shape
        bias = alibi if isinstance(alibi, AttnBias) else \
            AttnBias.from_reference_args(alibi, attention_mask, bsz, self.num_heads)
        qkv = F.linear(hidden_states, self.query_key_value.weight, self.query_key_value.bias)
        if isinstance(k_v_past, ops.StaticKV):
            # captured decode step (generation.py): device-side cache position / length
            q, k, v = F.split_packed(qkv, self.num_heads, F.LAYOUT_BLOOM)
            ctx, _ = ops.attn_fwd(q, k_v_past.k, k_v_past.v, self.inv_norm_factor, bias.causal, -ops.FLT_MAX,
                                  bias.kbias2, bias.first_valid, need_lse=False, seq_len_dev=k_v_past.len_dev,
                                  kv_new=(k, v))  # the kernel appends k, v itself
            out = F.linear(ctx, self.dense.weight, self.dense.bias, residual=residual)
            return out, k_v_past
        if k_v_past is None and torch.is_grad_enabled() and qkv.requires_grad:
            ctx = F.PackedAttentionFn.apply(qkv, self.num_heads, F.LAYOUT_BLOOM, self.inv_norm_factor,
                                            bias.causal, -ops.FLT_MAX, bias.kbias2, bias.first_valid, a_drop)
            _, k, v = F.split_packed(qkv.detach(), self.num_heads, F.LAYOUT_BLOOM)
        else:
            q, k, v = F.split_packed(qkv, self.num_heads, F.LAYOUT_BLOOM)
            if not torch.is_grad_enabled():
                # cache layout [b,h,t,d] (modeling_bloom.py:88-92), grown in place instead of concat
                k = ops.kv_cache_append(None if k_v_past is None else k_v_past[0], k)
                v = ops.kv_cache_append(None if k_v_past is None else k_v_past[1], v)
            elif k_v_past is not None:
                k = torch.cat((k_v_past[0], k), dim=-2)
                v = torch.cat((k_v_past[1], v), dim=-2)
            ctx = F.attention_cached(q, k, v, self.inv_norm_factor, bias.causal, -ops.FLT_MAX,
                                     bias.kbias2, bias.first_valid, a_drop)
        if h_on:
            out = F.dropout(F.linear(ctx, self.dense.weight, self.dense.bias, out_dtype=torch.float32),
                            self.hidden_dropout, True, residual=residual)
        else:
            out = F.linear(ctx, self.dense.weight, self.dense.bias, residual=residual)
        return out, (k, v)


class BloomMLP(torch.nn.Module):
    """modeling_bloom.py:243-271."""

    def __init__(self, config):
        super().__init__()
        hidden_size = config.hidden_size
        self.pretraining_tp = config.pretraining_tp
        self.slow_but_exact = config.slow_but_exact
        self.dense_h_to_4h = torch.nn.Linear(hidden_size, 4 * hidden_size)
        self.gelu_impl = BloomGelu()
        self.dense_4h_to_h = torch.nn.Linear(4 * hidden_size, hidden_size)
        self.hidden_dropout = config.hidden_dropout

    def forward(self, hidden_states, residual):
        h = F.linear(hidden_states, self.dense_h_to_4h.weight, self.dense_h_to_4h.bias, act=ops.ACT_GELU_TANH)
        if self.training and self.hidden_dropout > 0:  # modeling_bloom.py:269: residual + dropout(dense_4h_to_h(...))
            return F.dropout(F.linear(h, self.dense_4h_to_h.weight, self.dense_4h_to_h.bias, out_dtype=torch.float32),
                             self.hidden_dropout, True, residual=residual)
        return F.linear(h, self.dense_4h_to_h.weight, self.dense_4h_to_h.bias, residual=residual)


class BloomBlock(torch.nn.Module):
    """modeling_bloom.py:127-159 (pre-LN; optional residual-from-LayerNorm switch)."""

    def __init__(self, config):
        super(BloomBlock, self).__init__()
        hidden_size = config.hidden_size
        self.input_layernorm = LayerNorm(hidden_size, eps=config.layer_norm_epsilon)
        self.num_heads = config.n_head
        self.self_attention = BloomAttentionLayer(config)
        self.post_attention_layernorm = LayerNorm(hidden_size, eps=config.layer_norm_epsilon)
        self.mlp = BloomMLP(config)
        self.apply_residual_connection_post_layernorm = config.apply_residual_connection_post_layernorm
        self.hidden_dropout = config.hidden_dropout

    def _fused_spec(self):
        sa = self.self_attention
        return dict(ln1=self.input_layernorm, qkv=sa.query_key_value, proj=sa.dense,
                    ln2=self.post_attention_layernorm, fc1=self.mlp.dense_h_to_4h, fc2=self.mlp.dense_4h_to_h,
                    w_in_out=False, n_head=sa.num_heads, layout=F.LAYOUT_BLOOM, scale=sa.inv_norm_factor,
                    causal=True, causal_fill=-ops.FLT_MAX, act=ops.ACT_GELU_TANH)

    def forward(self, hidden_states, attention_mask, alibi, head_mask, k_v_past=None):
        cd = F.compute_dtype()
        post = self.apply_residual_connection_post_layernorm
        hidden_states = hidden_states if hidden_states.dtype == torch.float32 else hidden_states.float()
        if (isinstance(alibi, AttnBias) and k_v_past is None and not post and torch.is_grad_enabled()
                and hidden_states.requires_grad and alibi.causal and self.hidden_dropout == 0
                and self.self_attention.attention_dropout.p == 0):
            # training / full-sequence path: one fused autograd node for the whole block
            spec = self._fused_spec()
            out, k, v = F.PreLNBlockFn.apply(hidden_states, spec, alibi.kbias2, alibi.first_valid)
            return out, (k, v)
        if post:
            res1, ln1 = self.input_layernorm(hidden_states, out_dtype=torch.float32, out2_dtype=cd)
        else:
            ln1 = self.input_layernorm(hidden_states, out_dtype=cd)
            res1 = hidden_states
        att, k_v_past = self.self_attention(ln1, res1, attention_mask=attention_mask, alibi=alibi,
                                            head_mask=head_mask, k_v_past=k_v_past)
        if post:
            res2, ln2 = self.post_attention_layernorm(att, out_dtype=torch.float32, out2_dtype=cd)
        else:
            ln2 = self.post_attention_layernorm(att, out_dtype=cd)
            res2 = att
        return self.mlp(ln2, res2), k_v_past


class BloomModel(torch.nn.Module):
    """modeling_bloom.py:162-205."""

    def __init__(self, config):
        super(BloomModel, self).__init__()
        self.config = config
        self.num_heads = config.n_head
        self.embed_dim = config.hidden_size
        self.word_embeddings = torch.nn.Embedding(config.vocab_size, self.embed_dim)
        self.word_embeddings_layernorm = LayerNorm(self.embed_dim, eps=config.layer_norm_epsilon)
        self.blocks = torch.nn.ModuleList([BloomBlock(config) for _ in range(config.num_hidden_layers)])
        self.ln_f = LayerNorm(self.embed_dim, eps=config.layer_norm_epsilon)

    def forward(self, input_ids, attention_mask, head_mask, k_v_pasts=None):
        F.reject_head_mask(head_mask)
        if k_v_pasts is None:
            k_v_pasts = [None] * self.config.n_layer
        ln = self.word_embeddings_layernorm
        hidden_states = F.embedding_layer_norm([input_ids], [self.word_embeddings.weight], ln.weight, ln.bias, ln.eps)
        bias = attention_mask if isinstance(attention_mask, AttnBias) else \
            AttnBias.from_mask(attention_mask, self.num_heads, input_ids.shape[1])
        for i, block in enumerate(self.blocks):
            hidden_states, k_v_pasts[i] = block(hidden_states, attention_mask=None, alibi=bias,
                                                head_mask=None, k_v_past=k_v_pasts[i])
        return self.ln_f(hidden_states), k_v_pasts


class BloomForCausalLM(torch.nn.Module, GenerationMixin):
    """modeling_bloom.py:208-232."""

    def __init__(self, config):
        super(BloomForCausalLM, self).__init__()
        self.config = config
        self.bloom = BloomModel(config)
        self.lm_head = torch.nn.Linear(config.hidden_size, config.vocab_size, bias=False)

    def _tie_weight(self):
        self.lm_head.weight = self.bloom.word_embeddings.weight
        # two gradient contributions per step (lm_head wgrad first, embedding scatter last): the DDP
        # wrapper reduces this bucket only after both have been written
        self.lm_head.weight._ct_expected_writes = 2
        self.lm_head.weight._ct_sparse_second_write = True  # ... the second one being the token scatter

    _ct_graph_decode = True  # generation.py may replay the q_len = 1 step from a CUDA graph

    def _decode_static_mask(self, full_mask):
        """ALiBi + padding bias over the whole cache capacity, built once per generation; q_len = 1: no causal part."""
        b = AttnBias.from_mask(full_mask, self.bloom.num_heads, 1)
        return b

    def _decode_needs_positions(self):
        return False

    def _decode_graph_ok(self):
        return (self.config.hidden_size // self.config.n_head) in (32, 64, 128)  # csrc/attention.cu: attn_decode_kernel<D>

    def forward(self, input_ids, attention_mask=None, head_mask=None, k_v_pasts=None, labels=None, **kwargs):
        hidden_states, k_v_pasts = self.bloom(input_ids, attention_mask, head_mask, k_v_pasts)
        if labels is not None:
            loss, lm_logits = F.lm_head_loss(hidden_states, self.lm_head.weight, labels, shift=True)
            return (loss, lm_logits, hidden_states), k_v_pasts
        lm_logits = F.linear(hidden_states, self.lm_head.weight)
        return (lm_logits, hidden_states), k_v_pasts


class GeLUFunction(torch.autograd.Function):
    """modeling_bloom.py:275-285 — stand-alone tanh-GELU with the reference's hand-written backward
    (inside the MLP both directions are fused into GEMM epilogues instead)."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return bloom_gelu_forward(input)

    @staticmethod
    def backward(ctx, grad_output):
        return bloom_gelu_back(grad_output, ctx.saved_tensors)


class BloomGelu(torch.nn.Module):
    """modeling_bloom.py:289-305."""

    def forward(self, x):
        return GeLUFunction.apply(x) if self.training else bloom_gelu_forward(x)


def bloom_gelu_forward(x):
    """modeling_bloom.py:335-345: x*0.5*(1+tanh(0.79788456*x*(1+0.044715*x*x)))."""
    return ops.act_fwd(x, ops.ACT_GELU_TANH)


def bloom_gelu_back(g, x):
    """modeling_bloom.py:348-363 (x arrives as the 1-tuple of saved tensors)."""
    x = x[0] if isinstance(x, (tuple, list)) else x
    return ops.act_bwd(g, x, ops.ACT_GELU_TANH, out_dtype=g.dtype)
