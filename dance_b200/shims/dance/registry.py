"""Scoped preprocessor registry (reference dance/registry.py:190-233): ``register_preprocessor(*scope)`` files a transform
class under a dotted scope; ``REGISTERED`` is the scope → {name: class} table."""
REGISTERED = {}


def register_preprocessor(*scope: str):
    def deco(cls):
        REGISTERED.setdefault(".".join(("preprocessor", *scope)), {})[cls.__name__] = cls
        return cls
    return deco


def registered(*scope: str):
    return dict(REGISTERED.get(".".join(("preprocessor", *scope)), {}))
