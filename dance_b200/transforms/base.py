"""``BaseTransform`` — same contract as the reference (dance/transforms/base.py:12-49): ctor kwargs ``out``
and ``log_level``, ``name``, ``_DISPLAY_ATTRS``-driven ``__repr__`` and the md5 ``hexdigest`` built from it
(the dataset cache key, datasets/base.py:129-133)."""
from __future__ import annotations

import hashlib
import logging
from abc import ABC, abstractmethod
from typing import Optional, Tuple

logger = logging.getLogger("dance_b200")


class BaseTransform(ABC):
    _DISPLAY_ATTRS: Tuple[str, ...] = ()

    def __init__(self, out: Optional[str] = None, log_level="WARNING"):
        self.out = out or self.name
        self.logger = logger.getChild(self.name)
        self.logger.setLevel(log_level)
        self.log_level = log_level

    @property
    def name(self) -> str:
        return self.__class__.__name__

    def hexdigest(self) -> str:
        return hashlib.md5(repr(self).encode()).hexdigest()

    def __repr__(self) -> str:
        display_attrs_str = ", ".join(f"{i}={getattr(self, i)!r}" for i in self._DISPLAY_ATTRS)
        return f"{self.name}({display_attrs_str})"

    @abstractmethod
    def __call__(self, data):
        raise NotImplementedError
