"""Name -> class plug-in registry.

Mirrors the reference's registry contract (``step_recognition/utils/registry.py:6-19``,
``model/model_builder.py:5-9``): ``REG.register("name")`` is a class decorator,
``REG.register("name", obj)`` registers directly, duplicate names assert, and the
builders index the registry with a key of the flat config dict.
"""
from __future__ import annotations


class Registry(dict):
    def register(self, name, obj=None):
        def _add(o):
            assert name not in self, f"'{name}' is already registered"
            self[name] = o
            return o

        if obj is not None:
            _add(obj)
            return None
        return _add


META_ARCHITECTURES = Registry()
EVAL = Registry()


def build_model(cfg, device=None):
    """``model/model_builder.py:7-9``: ``META_ARCHITECTURES[cfg["model"]](cfg).to(device)``."""
    model = META_ARCHITECTURES[cfg["model"]](cfg)
    return model.to(device)


def build_eval(cfg):
    """``trainer/eval_builder.py:9-11``: ``EVAL[cfg["task"]](cfg)``."""
    return EVAL[cfg["task"]](cfg)
