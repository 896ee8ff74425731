"""Broadcasting dataclass base for RayBundle / RaySamples / Frustums.

Behavioural mirror of NS/utils/tensor_dataclass.py:35-332 (the objects that cross the drop-in boundary,
SURVEY.md 8b): every tensor field has shape ``[*batch, last_dim]``; on construction all fields are broadcast
to the common batch shape; indexing / reshape / flatten / to apply to the batch dims only.  Non-tensor fields
(e.g. ``spacing_to_euclidean_fn``) are carried through untouched.  Written from the behaviour, not the code.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Tuple

import numpy as np
import torch


class TensorDataclass:
    _shape: tuple = ()

    # ---- construction ---------------------------------------------------------------------------
    def __post_init__(self) -> None:
        if not dataclasses.is_dataclass(self):
            raise TypeError("TensorDataclass must be a dataclass")
        shapes = []
        self._visit(lambda k, v: shapes.append(v.shape[:-1]) if isinstance(v, torch.Tensor) else shapes.append(v.shape))
        if not shapes:
            raise ValueError("TensorDataclass must have at least one tensor")
        batch = torch.broadcast_shapes(*shapes)

        def bcast(k, v):
            if isinstance(v, torch.Tensor):
                return v.broadcast_to((*batch, v.shape[-1]))
            return v.broadcast_to(batch)

        self._map_inplace(bcast)
        object.__setattr__(self, "_shape", tuple(batch))

    def _items(self):
        for f in dataclasses.fields(self):
            yield f.name, getattr(self, f.name)

    def _visit(self, fn: Callable[[str, Any], None]) -> None:
        def rec(k, v):
            if isinstance(v, (torch.Tensor, TensorDataclass)):
                fn(k, v)
            elif isinstance(v, dict):
                for kk, vv in v.items():
                    rec(kk, vv)

        for k, v in self._items():
            rec(k, v)

    @staticmethod
    def _apply_value(v, fn):
        if isinstance(v, (torch.Tensor, TensorDataclass)):
            return fn(None, v)
        if isinstance(v, dict):
            return {kk: TensorDataclass._apply_value(vv, fn) for kk, vv in v.items()}
        return v

    def _map_inplace(self, fn) -> None:
        for k, v in list(self._items()):
            object.__setattr__(self, k, self._apply_value(v, fn))

    def _map(self, tensor_fn, dataclass_fn=None):
        dataclass_fn = dataclass_fn or tensor_fn

        def fn(_, v):
            return tensor_fn(v) if isinstance(v, torch.Tensor) else dataclass_fn(v)

        new_fields = {k: self._apply_value(v, fn) for k, v in self._items()}
        return dataclasses.replace(self, **new_fields)

    # ---- shape protocol --------------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    @property
    def size(self) -> int:
        return int(np.prod(self._shape)) if len(self._shape) else 1

    @property
    def ndim(self) -> int:
        return len(self._shape)

    def __len__(self) -> int:
        if len(self._shape) == 0:
            raise TypeError("len() of a 0-d tensor")
        return self._shape[0]

    def __bool__(self) -> bool:
        if len(self) == 0:
            raise ValueError(f"The truth value of {self.__class__.__name__} when `len(x) == 0` is ambiguous.")
        return True

    def __setitem__(self, indices, value):
        raise RuntimeError("Index assignment is not supported for TensorDataclass")

    def __getitem__(self, indices):
        if isinstance(indices, torch.Tensor):
            return self._map(lambda x: x[indices])
        if isinstance(indices, (int, slice, type(Ellipsis), list)):
            indices = (indices,)
        assert isinstance(indices, tuple)
        return self._map(lambda x: x[indices + (slice(None),)], lambda x: x[indices])

    def reshape(self, shape: Tuple[int, ...]):
        if isinstance(shape, int):
            shape = (shape,)
        return self._map(lambda x: x.reshape((*shape, x.shape[-1])), lambda x: x.reshape(shape))

    def flatten(self):
        return self.reshape((-1,))

    def broadcast_to(self, shape):
        return self._map(lambda x: x.broadcast_to((*shape, x.shape[-1])), lambda x: x.broadcast_to(shape))

    def to(self, device):
        return self._map(lambda x: x.to(device))
