"""Host-side mirror of the MODEL half of CleanTransformer/models/modeling_bert.py (lines 229-333)
over the sm_100a kernels. The tokenizer half of that file (BasicTokenizer / BertTokenizer, CPU
string processing) is outside the hot path and is not rebuilt.

Per block: three projection GEMMs -> attention kernel (additive (1-m)*-1e4 key mask,
modeling_bert.py:303-304) -> attention_post GEMM + bias + residual epilogue -> LayerNorm (f32 + bf16
copies) -> FFN GEMM + erf-GELU epilogue -> GEMM + bias + residual -> LayerNorm.
"""
import torch

from .. import functional as F
from .. import ops
from ..transformer import AttentionLayer, LayerNorm
from .modeling_gpt import _ActFn


class BertConfig():
    """modeling_bert.py:17-47."""

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.hidden_act = hidden_act
        self.intermediate_size = intermediate_size
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.pad_token_id = pad_token_id
        for k, v in kwargs.items():
            setattr(self, k, v)


class _Gelu(torch.nn.Module):
    def forward(self, x):
        return _ActFn.apply(x, ops.ACT_GELU_ERF)


class _Relu(torch.nn.Module):
    def forward(self, x):
        return _ActFn.apply(x, ops.ACT_RELU)


ACT2FN = {"gelu": _Gelu, "relu": _Relu}
_ACT_ID = {"gelu": ops.ACT_GELU_ERF, "relu": ops.ACT_RELU}


def _drop_on(mod):
    return mod.training and mod.p > 0


class BertTransformerBlock(torch.nn.Module):
    """modeling_bert.py:232-264 (post-LN)."""

    def __init__(self, config):
        super(BertTransformerBlock, self).__init__()
        self.config = config
        self.attention = AttentionLayer(config)
        self.attention_post = torch.nn.Sequential(
            torch.nn.Linear(config.hidden_size, config.hidden_size),
            torch.nn.Dropout(config.hidden_dropout_prob)
        )
        self.norm1 = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.ffw = torch.nn.Sequential(
            torch.nn.Linear(config.hidden_size, config.intermediate_size),
            ACT2FN[config.hidden_act](),
            torch.nn.Linear(config.intermediate_size, config.hidden_size),
        )
        self.dropout = torch.nn.Dropout(config.hidden_dropout_prob)
        self.norm2 = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states, attention_mask=None):
        cd = F.compute_dtype()
        hidden_states = hidden_states if hidden_states.dtype == torch.float32 else hidden_states.float()
        ctx = self.attention(hidden_states, attention_mask)
        post = self.attention_post[0]
        if _drop_on(self.attention_post[1]):  # modeling_bert.py:253-256: dropout(linear) + hidden, then LayerNorm
            a = F.linear(ctx, post.weight, post.bias, out_dtype=torch.float32)
            s1 = F.dropout(a, self.attention_post[1].p, True, residual=hidden_states)
        else:
            s1 = F.linear(ctx, post.weight, post.bias, residual=hidden_states)
        n1, n1_low = self.norm1(s1, out_dtype=torch.float32, out2_dtype=cd)
        h = F.linear(n1_low, self.ffw[0].weight, self.ffw[0].bias, act=_ACT_ID[self.config.hidden_act])
        if _drop_on(self.dropout):  # modeling_bert.py:259-262
            f = F.linear(h, self.ffw[2].weight, self.ffw[2].bias, out_dtype=torch.float32)
            s2 = F.dropout(f, self.dropout.p, True, residual=n1)
        else:
            s2 = F.linear(h, self.ffw[2].weight, self.ffw[2].bias, residual=n1)
        return self.norm2(s2)


class BertModel(torch.nn.Module):
    """modeling_bert.py:267-312."""

    def __init__(self, config):
        super(BertModel, self).__init__()
        self.config = config
        self.word_embeddings = torch.nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = torch.nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.segment_embeddings = torch.nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.embedding_post = torch.nn.Sequential(
            LayerNorm(config.hidden_size, eps=config.layer_norm_eps),
            torch.nn.Dropout(config.hidden_dropout_prob)
        )
        self.blocks = torch.nn.ModuleList([BertTransformerBlock(config) for _ in range(config.num_hidden_layers)])
        self.pooler = torch.nn.Sequential(
            torch.nn.Linear(config.hidden_size, config.hidden_size),
            torch.nn.Tanh()
        )

    def forward(self, input_ids=None, attention_mask=None, segment_ids=None, position_ids=None):
        if position_ids is None:
            # the reference builds this on the CPU (modeling_bert.py:294-295) and would then fail on a
            # GPU; build it on the input's device instead
            position_ids = torch.arange(input_ids.shape[1], dtype=torch.long, device=input_ids.device)
        ln = self.embedding_post[0]
        hidden_states = F.embedding_layer_norm(
            [input_ids, segment_ids, position_ids[None, :] if position_ids.dim() == 1 else position_ids],
            [self.word_embeddings.weight, self.segment_embeddings.weight, self.position_embeddings.weight],
            ln.weight, ln.bias, ln.eps, padding_idx0=0)
        hidden_states = F.dropout(hidden_states, self.embedding_post[1].p, _drop_on(self.embedding_post[1]))
        kb = None
        if attention_mask is not None:
            kb2, _ = ops.attn_mask_prep(attention_mask, self.config.num_attention_heads, ops.MASK_BERT)
            kb = _KeyBias(kb2)
        for block in self.blocks:
            hidden_states = block(hidden_states, kb)
        first_token_tensor = hidden_states[:, 0]
        pooled = F.linear(first_token_tensor, self.pooler[0].weight, self.pooler[0].bias,
                          act=ops.ACT_TANH, out_dtype=torch.float32)
        return (hidden_states, pooled)


class _KeyBias:
    """Pre-built per-key bias in the log2 domain; transformer.AttentionLayer accepts it directly."""

    def __init__(self, kbias2):
        self.kbias2 = kbias2


class BertForSequenceClassification(torch.nn.Module):
    """modeling_bert.py:315-333 (returns the logits tensor; the reference has no loss, :332)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config)
        self.drop = torch.nn.Dropout(config.hidden_dropout_prob)
        self.classifier = torch.nn.Linear(config.hidden_size, config.num_labels)

    def forward(self, input_ids=None, attention_mask=None, segment_ids=None, position_ids=None):
        hidden_states, pooled_output = self.bert(input_ids, attention_mask, segment_ids, position_ids)
        pooled_output = F.dropout(pooled_output, self.drop.p, _drop_on(self.drop))
        return F.linear(pooled_output, self.classifier.weight, self.classifier.bias, out_dtype=torch.float32)


class BertTokenizer:
    """examples/inference_bert.py:12,65-67 constructs `BertTokenizer(vocab_file=...)` from this module and
    reads `input_ids / attention_mask / segment_ids` from `encode_plus`. The reference's own WordPiece
    implementation (modeling_bert.py:50-226) is CPU string processing and out of scope; this adapter
    delegates to `transformers.BertTokenizer` (which the reference asserts equality with,
    modeling_bert.py:336-372) and renames `token_type_ids`."""

    def __init__(self, vocab_file, do_lower_case=True, **kwargs):
        import transformers
        self._tok = transformers.BertTokenizer(vocab_file=vocab_file, do_lower_case=do_lower_case, **kwargs)

    def tokenize(self, text):
        return self._tok.tokenize(text)

    def convert_tokens_to_ids(self, tokens):
        return self._tok.convert_tokens_to_ids(tokens)

    def encode_plus(self, text, text_pair=None, padding=False, truncation=False, max_length=None, **kwargs):
        enc = self._tok(text, text_pair, padding=padding, truncation=truncation, max_length=max_length, **kwargs)
        return {"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"],
                "segment_ids": enc["token_type_ids"]}
