"""Stub for the absent third-party `einconv` package (TEST INFRASTRUCTURE ONLY).

Only needed so that `import curvlinops` (the reference, /root/reference) succeeds when
golden vectors are generated in the build container.  KFAC-reduce on Conv2d is the only
consumer (reference curvlinops/kfac_utils.py:173-177) and is off the hot path.
"""


def index_pattern(*args, **kwargs):
    raise NotImplementedError("einconv stub: not available in this container")
