"""Minimal stand-in for ``ml_collections.ConfigDict`` (not installed here): attribute + item access,
``initial_dictionary=``, ``.get``, ``**`` unpacking.  Enough for the reference's ``configs/*.py`` files,
which are loaded unchanged through ``configs.load_config_file``."""
from __future__ import annotations


class ConfigDict(dict):
    def __init__(self, initial_dictionary=None, **kwargs):
        super().__init__()
        for k, v in dict(initial_dictionary or {}, **kwargs).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}


class FrozenConfigDict(ConfigDict):
    def __init__(self, cfg):
        super().__init__(cfg)
        object.__setattr__(self, "_frozen", True)

    def __setitem__(self, k, v):
        if getattr(self, "_frozen", False):
            raise AttributeError("FrozenConfigDict is immutable")
        super().__setitem__(k, v)
