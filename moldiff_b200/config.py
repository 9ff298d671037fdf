"""YAML config surface.  The reference reads configs through `EasyDict` (utils/misc.py:22-24) and
accesses them both by attribute (`config.diff.time_dim`) and by `**`-splat (`**config.denoiser`);
`AttrDict` gives the same two behaviours without the third-party dependency."""
from __future__ import annotations

import os

import yaml

CONFIG_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs")


class AttrDict(dict):
    def __init__(self, mapping=None, **kw):
        super().__init__()
        for k, v in dict(mapping or {}, **kw).items():
            self[k] = v

    @staticmethod
    def _lift(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._lift(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._lift(v))

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


def load_config(path):
    with open(path, "r") as f:
        return AttrDict(yaml.safe_load(f))


def builtin_config(name):
    """e.g. builtin_config('train/train_MolDiff.yml')"""
    return load_config(os.path.join(CONFIG_ROOT, name))
