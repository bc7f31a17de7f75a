"""Minimal stand-in for ``torch_geometric.data.Data`` (PyG is not a dependency): an attribute
container whose node-level tensors are recognised by ``size(0) == num_nodes`` exactly as the
reference's collate does (loader.py:187-190)."""
from typing import Iterator, Tuple

from torch import Tensor

from .sparse import SparseTensor


class Data:
    def __init__(self, **kwargs):
        self._store = {}
        for k, v in kwargs.items():
            self[k] = v

    def __getattr__(self, key):
        store = self.__dict__.get("_store", {})
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key == "_store":
            object.__setattr__(self, key, value)
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self._store

    def __iter__(self) -> Iterator[Tuple[str, object]]:
        return iter(list(self._store.items()))

    def keys(self):
        return list(self._store.keys())

    @property
    def num_nodes(self) -> int:
        if "x" in self._store:
            return self._store["x"].size(0)
        if "adj_t" in self._store:
            return self._store["adj_t"].size(0)
        if "y" in self._store:
            return self._store["y"].size(0)
        raise AttributeError("num_nodes")

    @property
    def num_edges(self) -> int:
        return self._store["adj_t"].nnz() if "adj_t" in self._store else 0

    def __copy__(self) -> "Data":
        out = self.__class__()
        out._store = dict(self._store)
        return out

    def to(self, device, non_blocking: bool = False) -> "Data":
        out = self.__class__()
        for k, v in self:
            if isinstance(v, (Tensor, SparseTensor)):
                v = v.to(device, non_blocking=non_blocking)
            out[k] = v
        return out

    def pin_memory(self) -> "Data":
        """Host-resident copy in page-locked memory (the reference's layout; the GPU collate reads it
        through UVA)."""
        out = self.__class__()
        for k, v in self:
            if isinstance(v, Tensor):
                v = v.cpu().pin_memory()
            elif isinstance(v, SparseTensor):
                v = v.pin_memory()
            out[k] = v
        return out

    def __repr__(self):
        parts = []
        for k, v in self:
            parts.append(f"{k}={list(v.shape) if isinstance(v, Tensor) else v}")
        return f"Data({', '.join(parts)})"
