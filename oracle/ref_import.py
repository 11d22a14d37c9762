"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules as a checker.

`/root/reference/{elvis,utils}.py` import heavyweight third-party packages at module
scope (lpips, skimage, fvmd, instantir, pytorch_msssim -- elvis.py:29-39, utils.py:20)
that are absent from this image.  None of them is touched by the hot-path functions,
so they are replaced by MagicMock stubs before the module body executes.

This module only works inside the build container (the reference tree is not shipped
to the GPU box).  It is used by `tests/golden/gen_golden.py` to freeze golden vectors
and by the `-m "not gpu"` oracle tests (skipped when the tree is absent).  Nothing in
`elvis_b200/` may import it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("ELVIS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "lpips", "skimage", "skimage.metrics", "pytorch_msssim", "instantir",
    "fvmd", "fvmd.datasets", "fvmd.datasets.video_datasets", "fvmd.keypoint_tracking",
    "fvmd.extract_motion_features", "fvmd.frechet_distance",
]
_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "elvis.py"))


def load(name: str):
    """Return the reference module `elvis` or `utils` (presley.py runs an experiment at
    import time -- presley.py:164-210 -- and cannot be imported)."""
    if name not in ("elvis", "utils"):
        raise ValueError(name)
    if name in _cache:
        return _cache[name]
    if not available():
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    for s in _STUBS:
        sys.modules.setdefault(s, MagicMock())
    spec = importlib.util.spec_from_file_location(f"_elvis_reference_{name}",
                                                  os.path.join(REFERENCE_ROOT, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def load_presley_functions(*names: str) -> dict:
    """presley.py cannot be imported (it runs its experiment at import time), but its hot-path
    functions are self-contained: compile the UNMODIFIED definitions of `names` out of the file's AST
    into a fresh namespace holding numpy / cv2 / typing.  Nothing is copied into this repository."""
    import ast
    import typing
    import numpy as np
    key = ("presley",) + names
    if key in _cache:
        return _cache[key]
    path = os.path.join(REFERENCE_ROOT, "presley.py")
    tree = ast.parse(open(path).read(), path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    if {n.name for n in wanted} != set(names):
        raise KeyError(f"presley.py lacks {set(names) - {n.name for n in wanted}}")
    ns: dict = {"np": np, "List": typing.List, "Tuple": typing.Tuple, "Any": typing.Any, "Callable": typing.Callable,
                "Optional": typing.Optional, "Dict": typing.Dict}
    try:
        import cv2
        ns["cv2"] = cv2
    except ImportError:
        pass
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), ns)
    _cache[key] = {n: ns[n] for n in names}
    return _cache[key]
