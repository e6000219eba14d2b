"""Host-side mirror of CleanTransformer/transformer.py over the sm_100a kernels.

Same class names, constructor arguments, attribute / parameter names (so state_dicts interchange
with the reference) and forward signatures; the arithmetic is done by libct_b200.so:
  AttentionLayer  -> 3 tcgen05 GEMMs (+bias) and the fused flash-style attention kernel
  LayerNorm       -> ct_layernorm_fwd/bwd
  TransformerBlock-> the above + GEMM epilogues for ReLU and the residual adds
`MultiHeadAttention` is exported as an alias of AttentionLayer (BASELINE.json's north_star uses
that name; the reference only has it as a README heading).

Dropout: the reference applies torch.nn.Dropout modules; they are kept as modules (p, train/eval state) and are the
identity in eval mode or with p = 0. With p > 0 in training mode the attention kernel drops the probabilities itself
(forward and backward regenerate the mask from a counter) and a hidden-state site runs ct_dropout, fused with the
residual add that follows it, instead of the GEMM's residual epilogue (functional.dropout).
"""
import math

import torch

from . import functional as F
from . import ops


class LayerNorm(torch.nn.Module):
    """transformer.py:61-89."""

    def __init__(self, normalized_shape, eps=1e-5):
        super(LayerNorm, self).__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        self.normalized_shape, self.eps = normalized_shape, eps
        self.weight = torch.nn.Parameter(torch.ones(normalized_shape))
        self.bias = torch.nn.Parameter(torch.zeros(normalized_shape))

    def forward(self, x, out_dtype=None, out2_dtype=None):
        return F.layer_norm(x, self.weight, self.bias, self.eps, out_dtype, out2_dtype)


def _dropout_active(mod):
    return mod is not None and mod.training and mod.p > 0


class AttentionLayer(torch.nn.Module):
    """transformer.py:12-58: separate q/k/v projections, softmax(QK^T/sqrt(d) + mask) V, no
    out-projection. attention_mask is the additive mask of the reference ([b,1,1,s] or broadcastable);
    a per-key mask is folded into the attention kernel."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        assert config.hidden_size % config.num_attention_heads == 0
        self.dim, self.m_head = config.hidden_size, config.num_attention_heads
        self.q_linear = torch.nn.Linear(config.hidden_size, config.hidden_size)
        self.k_linear = torch.nn.Linear(config.hidden_size, config.hidden_size)
        self.v_linear = torch.nn.Linear(config.hidden_size, config.hidden_size)
        self.dropout = torch.nn.Dropout(config.attention_probs_dropout_prob)

    @staticmethod
    def key_bias_from_additive(attention_mask, bsz, seq):
        """[b,1,1,s] additive mask (modeling_bert.py:303-304) -> kbias2 [b,1,s] in the log2 domain."""
        if attention_mask is None:
            return None
        if hasattr(attention_mask, "kbias2"):  # already prepared by the model (ops.attn_mask_prep)
            return attention_mask.kbias2
        m = attention_mask
        if m.dim() == 4 and m.shape[1] == 1 and m.shape[2] == 1:
            m = m.reshape(m.shape[0], 1, m.shape[3])
        elif m.dim() == 2:
            m = m[:, None, :]
        else:
            raise NotImplementedError("only per-key additive masks ([b,1,1,s]) are supported by the fused kernel")
        m = m.expand(bsz, 1, seq).float() * 1.4426950408889634
        return m.contiguous()

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        drop = F.next_dropout(self.dropout.p) if _dropout_active(self.dropout) else None  # transformer.py:47-50
        F.reject_head_mask(head_mask)
        b, s, _ = hidden_states.shape
        q = F.linear(hidden_states, self.q_linear.weight, self.q_linear.bias)
        k = F.linear(hidden_states, self.k_linear.weight, self.k_linear.bias)
        v = F.linear(hidden_states, self.v_linear.weight, self.v_linear.bias)
        kb = self.key_bias_from_additive(attention_mask, b, s)
        scale = 1.0 / math.sqrt(self.dim / self.m_head)
        return F.SeparateAttentionFn.apply(q, k, v, self.m_head, scale, False, -ops.FLT_MAX, kb, None, drop)


MultiHeadAttention = AttentionLayer


class TransformerBlock(torch.nn.Module):
    """transformer.py:92-121 (post-LN, ReLU FFN)."""

    def __init__(self, config):
        super(TransformerBlock, self).__init__()
        self.config = config
        self.attention = AttentionLayer(config)
        self.ffw = torch.nn.Sequential(
            torch.nn.Linear(config.hidden_size, config.hidden_size * 4),
            torch.nn.ReLU(),
            torch.nn.Linear(config.hidden_size * 4, config.hidden_size)
        )
        self.norm1 = LayerNorm(config.hidden_size, config.layer_norm_epsilong)
        self.norm2 = LayerNorm(config.hidden_size, config.layer_norm_epsilong)
        self.dropout = torch.nn.Dropout(config.hidden_dropout_prob)

    def forward(self, x):
        att_out = self.attention(x)
        on = _dropout_active(self.dropout)
        # x + dropout(att): the residual add has no GEMM to ride on here (no out-projection)
        add_norm_out = self.norm1(F.dropout(att_out, self.dropout.p, on, residual=x.float(), out_dtype=torch.float32))
        h = F.linear(add_norm_out, self.ffw[0].weight, self.ffw[0].bias, act=ops.ACT_RELU)
        if on:
            ffw_out = F.linear(h, self.ffw[2].weight, self.ffw[2].bias, out_dtype=torch.float32)
            return self.norm2(F.dropout(ffw_out, self.dropout.p, True, residual=add_norm_out))
        summed = F.linear(h, self.ffw[2].weight, self.ffw[2].bias, residual=add_norm_out)
        return self.norm2(summed)


class ExampleConfig():
    """transformer.py:124-131."""

    def __init__(self):
        self.num_attention_heads = 3
        self.layer_norm_epsilong = 1e-5
        self.resid_pdrop = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.hidden_size = 12
        self.hidden_dropout_prob = 0.1
