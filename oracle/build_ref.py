"""Recipe for ``oracle/_ref``: a runnable copy of the UNMODIFIED reference package.  TEST / BENCH INFRASTRUCTURE.

    python oracle/build_ref.py            # /root/reference/curvlinops -> oracle/_ref/curvlinops (byte-identical)

The reference (f-dangel/curvlinops) is pure Python, so "building" it is a copy of its package directory from where
it lies under ``/root/reference`` plus the two stub packages of ``oracle/stubs`` for third-party imports that are
absent from this image and off the hot path (``einconv``: KFAC-reduce patches, ``curvlinops/kfac_utils.py:10-11``;
``linear_operator``: GPyTorch's ``linear_cg``, ``curvlinops/inverse.py:7``).  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box, where ``bench.py --impl
reference`` times the reference's own CPU path and ``bench.py`` times it on the B200 through torch CUDA (the
"library kernels on the same box" bar).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline
legs may import it; the product path never does.

A pip install of the reference is not possible here (its build backend needs ``setuptools_scm``, absent, and there
is no network); the copy is what ``pip install --target`` would produce for a pure-Python package.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/curvlinops"
DST = os.path.join(HERE, "_ref")


def tree_digest(root: str) -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def build(force: bool = False) -> str | None:
    """Copy the reference package (when ``/root/reference`` exists) and return the path to put on ``sys.path``."""
    pkg = os.path.join(DST, "curvlinops")
    if os.path.isdir(SRC):
        if force or not os.path.isdir(pkg) or tree_digest(pkg) != tree_digest(SRC):
            shutil.rmtree(DST, ignore_errors=True)
            os.makedirs(DST)
            shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            for stub in ("einconv", "linear_operator"):
                shutil.copytree(os.path.join(HERE, "stubs", stub), os.path.join(DST, stub),
                                ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
                f.write(f"copied unmodified from {SRC} (sha256/16 of the .py tree: {tree_digest(SRC)}) + oracle/stubs\n")
    return DST if os.path.isdir(pkg) else None


def import_reference():
    """``import curvlinops`` from ``oracle/_ref`` (raises ImportError with the reason when it is not there)."""
    if not os.path.isdir(os.path.join(DST, "curvlinops")):
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import curvlinops

    return curvlinops


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
