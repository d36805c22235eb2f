"""Minimal stand-in for the slice of ``flax.linen`` the reference's modules rely on.

The reference's ``Flow`` and bijectors are FLAX modules (flow.py:16, bijectors.py:28): they
are configured by dataclass fields, create their variables lazily in ``init`` and read /
update them through ``apply(variables, ..., mutable=[...], method=...)``.  FLAX itself is not
available in this image, so this file re-creates exactly that surface — same call
signatures, same variable-tree naming (``params`` / ``batch_stats``; children named after
the attribute, list children ``<attr>_<i>``; ``BatchNorm_0``, ``Dense_j`` inside a compact
method) — so the variable pytree is interchangeable with the reference's
(SURVEY.md §8b).  It holds no compute.
"""
from __future__ import annotations

import contextlib
from typing import Any, Dict, Iterable, Optional, Sequence, Tuple, Union

import numpy as np

__all__ = ["Module", "Scope", "make_rng"]


def make_rng(key) -> np.random.Generator:
    """Accepts what callers pass as a PRNG key: an int seed, a numpy Generator, or an
    array-like (e.g. the two uint32 words of a ``jax.random.PRNGKey``)."""
    if isinstance(key, np.random.Generator):
        return key
    if key is None:
        return np.random.default_rng(0)
    arr = np.asarray(key)
    if arr.ndim == 0:
        return np.random.default_rng(int(arr))
    return np.random.default_rng([int(v) for v in arr.reshape(-1)])


def _copy_tree(t):
    return {k: _copy_tree(v) for k, v in t.items()} if isinstance(t, dict) else t


class Scope:
    """A cursor into the variable collections during ``init`` / ``apply``."""

    def __init__(self, root: Dict[str, dict], path: Tuple[str, ...] = (), mutable: Iterable[str] = (),
                 initializing: bool = False, rng: Optional[np.random.Generator] = None):
        self.root = root
        self.path = tuple(path)
        self.mutable = set(mutable)
        self.initializing = initializing
        self.rng = rng

    def child(self, name: str) -> "Scope":
        return Scope(self.root, self.path + (name,), self.mutable, self.initializing, self.rng)

    def _node(self, col: str, create: bool):
        node = self.root.get(col)
        if node is None:
            if not create:
                return None
            node = self.root.setdefault(col, {})
        for p in self.path:
            nxt = node.get(p)
            if nxt is None:
                if not create:
                    return None
                nxt = node.setdefault(p, {})
            node = nxt
        return node

    def has(self, col: str, name: str) -> bool:
        node = self._node(col, False)
        return node is not None and name in node

    def get(self, col: str, name: str):
        node = self._node(col, False)
        if node is None or name not in node:
            where = "/".join((col,) + self.path + (name,))
            raise KeyError(f'variable "{where}" not found; pass the variables returned by init()/train')
        return node[name]

    def subtree(self, col: str) -> dict:
        return self._node(col, False) or {}

    def is_mutable(self, col: str) -> bool:
        return self.initializing or col in self.mutable

    def put(self, col: str, name: str, value) -> None:
        if not self.is_mutable(col):
            where = "/".join((col,) + self.path + (name,))
            raise ValueError(f'cannot update variable "{where}": collection "{col}" is immutable '
                             f'(pass mutable=["{col}"])')
        self._node(col, True)[name] = value

    def variable(self, col: str, name: str, init_fn, *args):
        """flax ``self.variable``: create on first use (only while initializing)."""
        if not self.has(col, name):
            if not self.initializing:
                return self.get(col, name)  # raises KeyError with the path
            self.put(col, name, init_fn(*args))
        return self.get(col, name)


class Module:
    """Base of Flow and the bijectors: ``init`` / ``apply`` with FLAX semantics."""

    _scope: Optional[Scope] = None

    # -- binding ------------------------------------------------------------------------
    @contextlib.contextmanager
    def _bound(self, scope: Optional[Scope]):
        prev = self.__dict__.get("_scope")
        self.__dict__["_scope"] = scope
        try:
            yield self
        finally:
            self.__dict__["_scope"] = prev

    @property
    def scope(self) -> Scope:
        sc = self.__dict__.get("_scope")
        if sc is None:
            raise RuntimeError(f"{type(self).__name__} is not bound: call it through .init(...) or "
                               f".apply(variables, ...), as with a FLAX module")
        return sc

    def is_initializing(self) -> bool:
        return self.scope.initializing

    # -- public FLAX-like API -------------------------------------------------------------
    def init(self, rngs, *args, method=None, **kwargs) -> Dict[str, dict]:
        """Create the variable collections by running the module once (flax ``Module.init``)."""
        root: Dict[str, dict] = {}
        scope = Scope(root, (), (), True, make_rng(rngs))
        fn = self._resolve(method)
        with self._bound(scope):
            fn(*args, **kwargs)
        return {k: v for k, v in root.items() if v}

    def apply(self, variables, *args, method=None, mutable: Union[bool, str, Sequence[str]] = False,
              rngs=None, **kwargs):
        """Run a method with the given variables (flax ``Module.apply``).  With ``mutable``
        the call returns ``(output, updated_collections)`` and leaves ``variables`` untouched."""
        if mutable is True:
            mut = set(variables.keys()) | {"batch_stats"}
        elif not mutable:
            mut = set()
        elif isinstance(mutable, str):
            mut = {mutable}
        else:
            mut = set(mutable)
        root = {k: (_copy_tree(v) if k in mut else v) for k, v in variables.items()}
        scope = Scope(root, (), mut, False, make_rng(rngs) if rngs is not None else None)
        fn = self._resolve(method)
        with self._bound(scope):
            out = fn(*args, **kwargs)
        if mut:
            return out, {k: root.get(k, {}) for k in mut if k in root}
        return out

    def _resolve(self, method):
        if method is None:
            return self.__call__
        if isinstance(method, str):
            return getattr(self, method)
        # an unbound function such as Flow.sample
        name = getattr(method, "__name__", None)
        if name and hasattr(self, name):
            return getattr(self, name)
        return lambda *a, **k: method(self, *a, **k)
