"""Import the *unmodified* reference (``dmc/models``).  TEST INFRASTRUCTURE ONLY.

Where the reference is looked for, in order:

1. ``$DVC_REFERENCE_ROOT``
2. ``/root/reference``                 (build container; read-only)
3. ``<repo>/baseline/_ref``            (git-ignored staging directory written by
   ``tools/stage_reference.py``; it travels to the GPU box with the snapshot,
   so the stock ``DMC`` can be executed on the B200 -- ``tests/test_gpu_dropin.py``)

Nothing is copied into the repository history; the reference files are executed
where they lie.  ``compressai`` is absent everywhere, so a provider directory
is put on ``sys.path`` for the duration of the import:

* ``oracle/``  -> the eager restatement (``oracle/compressai``): this is the
  **stock** arm (the reference exactly as written + eager PyTorch ops);
* ``deepvideocodec_b200/compressai_shim`` -> the kernel-backed modules: the
  **patched** arm, completed by ``deepvideocodec_b200.patch(models)``.
"""
import importlib
import importlib.util
import os
import sys

_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_ORACLE_DIR)
STAGED_ROOT = os.path.join(_REPO_DIR, "baseline", "_ref")


def _has_reference(root):
    return bool(root) and os.path.isfile(os.path.join(root, "dmc", "models", "layers.py"))


def find_reference_root():
    for root in (os.environ.get("DVC_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if _has_reference(root):
            return root
    return None


REFERENCE_ROOT = find_reference_root() or "/root/reference"


def reference_available():
    return find_reference_root() is not None


def load_reference_models():
    """Returns the reference ``models`` package (``dmc/models``)."""
    root = find_reference_root()
    if root is None:
        raise FileNotFoundError("reference not found (DVC_REFERENCE_ROOT, /root/reference, "
                                f"{STAGED_ROOT})")
    try:
        importlib.import_module("compressai")
    except ModuleNotFoundError:
        if _ORACLE_DIR not in sys.path:
            sys.path.insert(0, _ORACLE_DIR)
    dmc_dir = os.path.join(root, "dmc")
    if dmc_dir not in sys.path:
        sys.path.insert(0, dmc_dir)
    return importlib.import_module("models")


def _is_compressai(name):
    return name == "compressai" or name.startswith("compressai.")


def load_reference_models_as(alias, compressai_dir):
    """Execute the reference's ``dmc/models`` package under the module name
    ``alias`` with ``import compressai`` resolved from ``compressai_dir``.
    Two independent copies of the reference classes (stock / patched) can then
    live in one process.  ``sys.modules['compressai*']`` is restored."""
    if alias in sys.modules:
        return sys.modules[alias]
    root = find_reference_root()
    if root is None:
        raise FileNotFoundError("reference not found")
    models_dir = os.path.join(root, "dmc", "models")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_compressai(k)}
    sys.path.insert(0, compressai_dir)
    importlib.invalidate_caches()
    try:
        spec = importlib.util.spec_from_file_location(
            alias, os.path.join(models_dir, "__init__.py"),
            submodule_search_locations=[models_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[alias] = pkg
        try:
            spec.loader.exec_module(pkg)
        except BaseException:
            for k in [k for k in sys.modules if k == alias or k.startswith(alias + ".")]:
                del sys.modules[k]
            raise
        pkg.__compressai_provider__ = sys.modules.get("compressai")
    finally:
        sys.path.remove(compressai_dir)
        for k in [k for k in sys.modules if _is_compressai(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
        importlib.invalidate_caches()
    return pkg


def load_stock_and_patched(**patch_kwargs):
    """``(stock_models, patched_models)``: the reference twice in one process.

    stock   = unmodified ``dmc/models`` over the eager CompressAI restatement;
    patched = unmodified ``dmc/models`` over ``deepvideocodec_b200``'s entropy
              modules with ``deepvideocodec_b200.patch`` applied (what a user
              gets from ``install_compressai_shim(); import models; patch(models)``).
    """
    import deepvideocodec_b200 as dvc
    patch_mod = importlib.import_module("deepvideocodec_b200.patch")
    stock = load_reference_models_as("dvc_ref_stock", _ORACLE_DIR)
    patched = load_reference_models_as("dvc_ref_patched", patch_mod._SHIM_DIR)
    dvc.patch(patched, **patch_kwargs)
    return stock, patched


def load_reference_train_ns(names):
    """Several top-level definitions of ``dmc/train.py`` executed into ONE scratch
    namespace (so ``RateDistortionLoss`` finds ``collect_likelihoods_list``);
    returned as a ``types.SimpleNamespace`` -- a stand-in for the ``train`` module
    that ``deepvideocodec_b200.patch(models, train_module=ns)`` can rebind."""
    import ast
    import math
    import types
    from collections import defaultdict
    from typing import List

    import torch
    path = os.path.join(find_reference_root() or REFERENCE_ROOT, "dmc", "train.py")
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in body}
    if missing:
        raise KeyError(sorted(missing))
    mod = types.ModuleType("dvc_ref_train_ns")
    mod.__dict__.update({"torch": torch, "nn": torch.nn, "optim": torch.optim, "math": math,
                         "defaultdict": defaultdict, "List": List})
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), mod.__dict__)
    return mod


def load_reference_train_fn(name):
    """Fetch one function from ``dmc/train.py`` *without importing the module*
    (importing it overwrites CUDA_VISIBLE_DEVICES, train.py:43, and pulls in
    torchvision/datasets).  The function's source lines are executed from the
    reference file in a scratch namespace."""
    import ast
    import math
    from collections import defaultdict

    import torch
    path = os.path.join(find_reference_root() or REFERENCE_ROOT, "dmc", "train.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "nn": torch.nn, "math": math, "defaultdict": defaultdict}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)
