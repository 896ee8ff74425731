"""Per-image appearance embedding (NS/field_components/embedding.py:27-59): ``nn.Embedding`` with ``mean`` over the table."""
from __future__ import annotations

import torch
from torch import nn


class Embedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.embedding = nn.Embedding(in_dim, out_dim)

    def get_out_dim(self) -> int:
        return self.out_dim

    def mean(self, dim=0) -> torch.Tensor:
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        return self.embedding(in_tensor)
