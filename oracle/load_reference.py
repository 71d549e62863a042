"""Import the *unmodified* reference from ``/root/reference`` (build container
only -- the GPU box has no copy).  TEST INFRASTRUCTURE ONLY.

``compressai`` is absent, so ``oracle/`` is put on ``sys.path`` first and the
reference's ``from compressai.entropy_models import ...`` resolves to the
oracle shim.  Nothing is copied; the reference files are executed where they
lie.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("DVC_REFERENCE_ROOT", "/root/reference")
_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dmc", "models", "layers.py"))


def load_reference_models():
    """Returns the reference ``models`` package (``dmc/models``)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    try:
        importlib.import_module("compressai")
    except ModuleNotFoundError:
        if _ORACLE_DIR not in sys.path:
            sys.path.insert(0, _ORACLE_DIR)
    dmc_dir = os.path.join(REFERENCE_ROOT, "dmc")
    if dmc_dir not in sys.path:
        sys.path.insert(0, dmc_dir)
    return importlib.import_module("models")


def load_reference_train_fn(name):
    """Fetch one function from ``dmc/train.py`` *without importing the module*
    (importing it overwrites CUDA_VISIBLE_DEVICES, train.py:43, and pulls in
    torchvision/datasets).  The function's source lines are executed from the
    reference file in a scratch namespace."""
    import ast
    import math
    from collections import defaultdict

    import torch
    path = os.path.join(REFERENCE_ROOT, "dmc", "train.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "math": math, "defaultdict": defaultdict}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)
