"""``Compose`` and ``SetConfig`` (reference dance/transforms/misc.py:15-71, 102-122)."""
from __future__ import annotations

import hashlib
from pprint import pformat
from typing import Any, Dict

from .base import BaseTransform


class Compose(BaseTransform):

    def __init__(self, *transforms, use_master_log_level: bool = True, **kwargs):
        super().__init__(**kwargs)
        failed = [t for t in transforms if not isinstance(t, BaseTransform)]
        if failed:
            raise TypeError("Expect all transform objects to be inherited from BaseTransform. The following "
                            f"(n={len(failed)}) have incorrect types:\n" + "\n".join(f"\t{i!r}: {type(i)!r}" for i in failed))
        self.transforms = transforms
        if use_master_log_level:
            for t in transforms:
                t.log_level = self.log_level
                t.logger.setLevel(self.log_level)

    def __repr__(self):
        return "Compose(\n  " + ",\n  ".join(map(repr, self.transforms)) + ",\n)"

    def __getitem__(self, idx: int, /):
        return self.transforms[idx]

    def hexdigest(self) -> str:
        return hashlib.md5("".join(t.hexdigest() for t in self.transforms).encode()).hexdigest()

    def __call__(self, data):
        self.logger.info(f"Applying composed transformations:\n{self!r}")
        for t in self.transforms:
            t(data)


class SetConfig(BaseTransform):
    _DISPLAY_ATTRS = ("config_dict", )

    def __init__(self, config_dict: Dict[str, Any], dummy_params=10, **kwargs):
        super().__init__(**kwargs)
        self.config_dict = config_dict

    def __call__(self, data):
        self.logger.info(f"Updating the dance data object config options:\n{pformat(self.config_dict)}")
        data.set_config_from_dict(self.config_dict)
