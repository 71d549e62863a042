"""Install the B200 hot path into the *unmodified* reference.

The reference has no plugin registry; its seam is Python name binding
(SURVEY.md 8b).  ``patch(models_pkg)`` rebinds, without touching any reference
file:

=====================================================  =========================================
reference name (dmc/models/...)                        replacement
=====================================================  =========================================
``video_model.flow_warp`` (:11-12), ``layers.flow_warp``
/ ``layers.torch_warp`` (SpyNet, layers.py:261)        ``layers.flow_warp``
``video_model.bilineardownsacling`` (:12)              ``layers.bilineardownsacling``
``video_model.quantize_ste`` (:10), ``utils.quantize_ste``  ``utils.quantize_ste``
``MotionContextModel.forward_dual_prior`` (:169)       ``context.forward_dual_prior``
``FrameContextModel.forward_dual_prior`` (:341)        ``context.forward_dual_prior``
``MotionContextModel.forward`` (:218)                  ``context.motion_context_forward``
``FrameContextModel.forward`` (:390)                   ``context.frame_context_forward``
``MotionContextModel.compress/decompress`` (:236/:255) ``context.motion_context_(de)compress``
``FrameContextModel.compress/decompress`` (:408/:429)  ``context.frame_context_(de)compress``
``DMC.motion_compensation`` (:497)                     fused 2-launch version below
``train.collect_likelihoods_list`` (train.py:74)       ``rate.collect_likelihoods_list``
=====================================================  =========================================

``install_compressai_shim()`` additionally makes ``import compressai`` resolve
to this package's entropy models when the real CompressAI is not installed, so
stock ``dmc/train.py`` / ``dmc/test.py`` import.
"""
import importlib
import os
import sys

from . import context, layers, rate, utils

_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compressai_shim")
_saved = []


def install_compressai_shim(force=False):
    """Put the ``compressai`` shim on ``sys.path`` unless the real package
    imports.  Returns True when the shim is (now) the provider."""
    if not force:
        try:
            mod = importlib.import_module("compressai")
            return getattr(mod, "__dvc_b200_shim__", False)
        except ModuleNotFoundError:
            pass
    if _SHIM_DIR not in sys.path:
        sys.path.insert(0, _SHIM_DIR)
    importlib.invalidate_caches()
    return True


def _motion_compensation(self, mv, dpb):
    """Drop-in for ``DMC.motion_compensation`` (video_model.py:497-506): the
    flow pyramid and the four warps are ONE launch."""
    ref_feature1, ref_feature2, ref_feature3 = self.multi_scale_feature_extractor(dpb)
    context1, context2, context3, warpframe = layers.motion_compensation_warps(
        dpb["x_ref"], ref_feature1, ref_feature2, ref_feature3, mv)
    context1, context2, context3 = self.context_fusion_net(context1, context2, context3)
    return context1, context2, context3, warpframe


def _set(obj, name, value):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def patch(models_pkg, train_module=None, fuse_context_models=True):
    """Rebind the hot-path names of the reference ``models`` package (the
    module object of ``dmc/models``).  Idempotent; ``unpatch()`` restores."""
    if _saved:
        return
    vm = sys.modules[models_pkg.__name__ + ".video_model"]
    ly = sys.modules[models_pkg.__name__ + ".layers"]
    ut = sys.modules[models_pkg.__name__ + ".utils"]
    _set(vm, "flow_warp", layers.flow_warp)
    _set(vm, "bilineardownsacling", layers.bilineardownsacling)
    _set(vm, "quantize_ste", utils.quantize_ste)
    _set(ly, "flow_warp", layers.flow_warp)
    _set(ly, "torch_warp", layers.flow_warp)
    _set(ly, "bilineardownsacling", layers.bilineardownsacling)
    _set(ut, "quantize_ste", utils.quantize_ste)
    for cls in (vm.MotionContextModel, vm.FrameContextModel):
        _set(cls, "forward_dual_prior", context.forward_dual_prior)
    if fuse_context_models:
        _set(vm.MotionContextModel, "forward", context.motion_context_forward)
        _set(vm.FrameContextModel, "forward", context.frame_context_forward)
        _set(vm.MotionContextModel, "compress", context.motion_context_compress)
        _set(vm.FrameContextModel, "compress", context.frame_context_compress)
        _set(vm.MotionContextModel, "decompress", context.motion_context_decompress)
        _set(vm.FrameContextModel, "decompress", context.frame_context_decompress)
    _set(vm.DMC, "motion_compensation", _motion_compensation)
    if train_module is not None:
        _set(train_module, "collect_likelihoods_list", rate.collect_likelihoods_list)


def unpatch():
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)
