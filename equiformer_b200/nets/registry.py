"""Model registry (same entry points as the reference's ``nets/registry.py``: ``register_model`` / ``model_entrypoint``)."""
from __future__ import annotations

_model_entrypoints = {}


def register_model(fn):
    _model_entrypoints[fn.__name__] = fn
    return fn


def model_entrypoint(model_name: str):
    return _model_entrypoints[model_name]


def is_model(model_name: str) -> bool:
    return model_name in _model_entrypoints


def list_models():
    return sorted(_model_entrypoints)
