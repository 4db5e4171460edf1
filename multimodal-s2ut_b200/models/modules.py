"""Parameter containers with fairseq-compatible ``state_dict`` names and initialisation.

These modules own the fp32 master parameters of the hot path; they have no eager ``forward``.
All arithmetic runs in the sm_100a kernels driven by ``engine.py``; there is deliberately no
PyTorch fallback path.  Names and shapes follow SURVEY.md §8(b) so that a checkpoint written by the
reference (fairseq ``S2TTransformerEncoder`` + mm_s2ut/models/mm_s2s_transformer.py:91-260 +
mm_s2ut/models/fuse.py:35-63) loads unchanged.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def Linear(in_features: int, out_features: int, bias: bool = True) -> nn.Linear:
    """xavier-uniform weight, zero bias (reference: mm_s2ut/models/fuse.py:18-23)."""
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.0)
    return m


class Conv1dSubsampler(nn.Module):
    """Parameters of fairseq ``Conv1dSubsampler(in, mid, out, kernel_sizes)``."""

    def __init__(self, in_channels: int, mid_channels: int, out_channels: int, kernel_sizes=(5, 5)):
        super().__init__()
        self.n_layers = len(kernel_sizes)
        self.conv_layers = nn.ModuleList(
            nn.Conv1d(in_channels if i == 0 else mid_channels // 2,
                      mid_channels if i < self.n_layers - 1 else out_channels * 2,
                      k, stride=2, padding=k // 2)
            for i, k in enumerate(kernel_sizes))


class SelfAttentionParams(nn.Module):
    """fairseq ``MultiheadAttention`` (self-attention): separate q/k/v/out Linear with bias."""

    def __init__(self, embed_dim: int, num_heads: int):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        g = 1 / math.sqrt(2)
        nn.init.xavier_uniform_(self.k_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.v_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.q_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class TransformerEncoderLayerParams(nn.Module):
    def __init__(self, embed_dim: int, ffn_dim: int, num_heads: int):
        super().__init__()
        self.self_attn = SelfAttentionParams(embed_dim, num_heads)
        self.self_attn_layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)
        self.fc1 = nn.Linear(embed_dim, ffn_dim)
        self.fc2 = nn.Linear(ffn_dim, embed_dim)
        self.final_layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)


class SinusoidalPositionalEmbedding(nn.Module):
    """Keeps fairseq's ``_float_tensor`` buffer for checkpoint compatibility; the table itself is
    regenerated on device by the engine (rows: [sin | cos], row ``padding_idx`` zero)."""

    def __init__(self, embedding_dim: int, padding_idx: int, init_size: int = 1024):
        super().__init__()
        self.embedding_dim, self.padding_idx, self.init_size = embedding_dim, padding_idx, init_size
        self.register_buffer("_float_tensor", torch.FloatTensor(1))

    @staticmethod
    def get_embedding(num_embeddings: int, dim: int, padding_idx: int) -> torch.Tensor:
        half = dim // 2
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
        e = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
        e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_embeddings, -1)
        if dim % 2 == 1:
            e = torch.cat([e, torch.zeros(num_embeddings, 1)], dim=1)
        e[padding_idx, :] = 0
        return e


class SelectiveAttentionParams(nn.Module):
    """Parameters of ``SelectiveAttention`` (reference: mm_s2ut/models/fuse.py:36-63)."""

    def __init__(self, qdim, kdim, vdim, attn_dim, intermediate_dim, output_dim, num_heads=1,
                 qkv_bias=True, attn_drop=0.0):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError("the mm_s2ut_transformer encoder builds SelectiveAttention with num_heads=1")
        self.num_heads, self.qdim, self.kdim, self.vdim = num_heads, qdim, kdim, vdim
        self.output_dim, self.intermediate_dim = output_dim, intermediate_dim
        self.qkhead_dim = attn_dim // num_heads
        self.vhead_dim = intermediate_dim // num_heads
        self.scale = self.qkhead_dim ** -0.5
        self.q_proj = Linear(qdim, attn_dim, bias=qkv_bias)
        self.k_proj = Linear(kdim, attn_dim, bias=qkv_bias)
        self.v_proj = Linear(vdim, intermediate_dim, bias=qkv_bias)
        self.attn_drop_p = attn_drop
        self.proj = Linear(intermediate_dim, output_dim)


class MultimodalAttentionParams(nn.MultiheadAttention):
    """Parameters of ``MultimodalAttention`` (reference: mm_s2ut/models/fuse.py:120-129): an
    ``nn.MultiheadAttention`` subclass, so its parameter names/initialisation are torch's own."""

    def forward(self, *a, **k):  # pragma: no cover - never called; the engine runs the kernels
        raise RuntimeError("MultimodalAttentionParams is a parameter container; use the encoder forward")


class Conv1dAdaptorParams(nn.Module):
    """Unused on this path but always constructed by the reference (mm_s2s_transformer.py:215-224),
    hence present in its checkpoints."""

    def __init__(self, in_dim=1024, out_dim=768, n_layers=3, kernel_size=3, stride=2, layernorm=True):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.Conv1d(in_dim if i == 0 else out_dim, out_dim * 2, kernel_size, stride=stride,
                      padding=kernel_size // 2) for i in range(n_layers))
        self.layernorm = nn.LayerNorm(in_dim) if layernorm else None
