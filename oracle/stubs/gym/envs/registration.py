import importlib

_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = _registry[id]
    mod, cls = entry_point.split(":")
    ctor = getattr(importlib.import_module(mod), cls)
    merged = dict(kw.get("kwargs", {}))
    merged.update(kwargs)
    return ctor(**merged)
