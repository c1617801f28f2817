"""Positional-embedding parameter containers with the reference's state-dict keys.

AxialPositionalEmbedding mirrors the third-party `axial_positional_embedding` package the reference imports
(dalle_bert.py:8,326; dalle_artv.py:8,141; parameters `weights_0..k`, shapes (1, 1.., n_i, ..1, dim), N(0,1)).
AxialPositionalEmbeddingList mirrors mmvid_pytorch/modules.py:8-53 (`module_list.{v}.weights_{0,1}`).
Instead of materialising a [b, n, dim] tensor per call, `table()` builds the [n, dim] position table once on
the GPU (mmvid_axial_table) for the fused embedding-gather kernel to add.
"""
import numpy as np
import torch
from torch import nn

from . import _lib, ops


class AxialPositionalEmbedding(nn.Module):
    def __init__(self, dim, axial_shape, axial_dims=None):
        super().__init__()
        self.dim = dim
        self.shape = tuple(int(s) for s in axial_shape)
        self.max_seq_len = int(np.prod(self.shape))
        for ind, n in enumerate(self.shape):
            ax_shape = [1] * len(self.shape)
            ax_shape[ind] = n
            self.register_parameter(f"weights_{ind}", nn.Parameter(torch.zeros(1, *ax_shape, dim).normal_(0, 1)))
        self._cache = None

    def axis_weights(self):
        return [getattr(self, f"weights_{i}") for i in range(len(self.shape))]

    def table(self):
        """float32 [max_seq_len, dim] on the parameters' device (cached until a parameter changes)."""
        ws = self.axis_weights()
        key = tuple((w.data_ptr(), w._version) for w in ws) + (_lib.weights_epoch(),)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                self._cache = (key, ops.axial_table([w.detach() for w in ws], self.shape))
        return self._cache[1]

    def forward(self, x):
        """Reference call signature: returns the position embedding broadcast to x's [b, t, dim]."""
        b, t, _ = x.shape
        return self.table()[:t].unsqueeze(0).expand(b, t, self.dim).to(x)


class AxialPositionalEmbeddingList(nn.Module):
    def __init__(self, dim=512, num=None, axial_shape=()):
        super().__init__()
        if num is None:
            num = axial_shape[0]
            axial_shape = axial_shape[1:]
        self.dim = dim
        self.num = num
        self.axial_shape = tuple(axial_shape)
        self.chunk_size = int(np.prod(axial_shape))
        self.seq_len = num * self.chunk_size
        self.module_list = nn.ModuleList([AxialPositionalEmbedding(dim, axial_shape=axial_shape) for _ in range(num)])

    def table(self):
        """float32 [num * chunk, dim]: per-frame tables concatenated (modules.py:45-52)."""
        return torch.cat([m.table() for m in self.module_list], dim=0)

    def forward(self, emb):
        b, t, _ = emb.shape
        if t > self.seq_len:
            raise NotImplementedError("insert_sep layouts are not on the hot path")
        return self.table()[:t].unsqueeze(0).expand(b, t, self.dim).to(emb)
