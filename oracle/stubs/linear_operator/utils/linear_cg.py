"""Stub for the absent `linear_operator` package (reference curvlinops/inverse.py:7)."""


def linear_cg(*args, **kwargs):
    raise NotImplementedError("linear_operator stub: not available in this container")
