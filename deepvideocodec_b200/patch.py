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
``DMC.motion_compensation`` (:497)                     one-launch warps below; in inference, with
                                                       ``fuse_warp_conv=True``, each warp is fused into
                                                       the ``conv{1,2,3}_out`` that consumes it (tcgen05)
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


def motion_compensation_fused(self, mv, dpb):
    """``DMC.motion_compensation`` + ``MultiScaleContextFusion.forward``
    (video_model.py:497-506, :49-66) with every context warp fused into the 3x3
    convolution that consumes it: the warped contexts are
    written once (the residual adds need them) and never re-read by
    ``conv{1,2,3}_out``; mv2 / mv3 are never materialised.  Same sub-modules,
    same parameters, same return structure; the three fused convs compute in
    TF32 like cuDNN's default, the rest of the net is the reference's own
    modules.  Inference only -- with autograd on, the unfused path runs."""
    import torch
    if torch.is_grad_enabled() or mv.size(2) % 4 or mv.size(3) % 4 or \
            any(getattr(self.context_fusion_net, n).out_channels != 64
                for n in ("conv1_out", "conv2_out", "conv3_out")):
        return _motion_compensation(self, mv, dpb)
    f1, f2, f3 = self.multi_scale_feature_extractor(dpb)
    net = self.context_fusion_net
    warpframe = layers.flow_warp(dpb["x_ref"], mv)
    c3, c3_pre = layers.warp_conv3x3(f3, mv, net.conv3_out.weight, net.conv3_out.bias,
                                     flow_downscale=2)
    c3_up = net.res_block3_up(net.conv3_up(c3))
    c3_out = net.res_block3_out(c3_pre)
    c2, c2_pre = layers.warp_conv3x3(f2, mv, net.conv2_out.weight, net.conv2_out.bias,
                                     extra=c3_up, flow_downscale=1)
    c2_up = net.res_block2_up(net.conv2_up(torch.cat((c3_up, c2), dim=1)))
    c2_out = net.res_block2_out(c2_pre)
    c1, c1_pre = layers.warp_conv3x3(f1, mv, net.conv1_out.weight, net.conv1_out.bias,
                                     extra=c2_up)
    c1_out = net.res_block1_out(c1_pre)
    return c1 + c1_out, c2 + c2_out, c3 + c3_out, warpframe


def _set(obj, name, value):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def _graphed_forward_inter(eager_inter):
    """``DMC.forward_inter`` replayed from one CUDA graph per P-frame in inference
    (``graph.GraphedInter``); the eager method everywhere else."""
    from .graph import GraphedInter

    def forward_inter(self, x_cur, dpb, motion_pretrain=False, frame_pretrain=False):
        g = self.__dict__.get("_dvc_graphed_inter")
        if g is None:
            g = GraphedInter(self, eager_inter=eager_inter)
            self.__dict__["_dvc_graphed_inter"] = g
        return g(x_cur, dpb, motion_pretrain, frame_pretrain)
    return forward_inter


def patch(models_pkg, train_module=None, fuse_context_models=True, fuse_warp_conv=False,
          graph_inter=False):
    """Rebind the hot-path names of the reference ``models`` package (the
    module object of ``dmc/models``).  Idempotent; ``unpatch()`` restores.

    ``graph_inter=True`` additionally replays ``DMC.forward_inter`` from one CUDA
    graph per P-frame whenever the model is in eval mode under ``no_grad``
    (2.9x at 256x256, 1.09x at 1080p on the reference's own
    model)."""
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
    _set(vm.DMC, "motion_compensation",
         motion_compensation_fused if fuse_warp_conv else _motion_compensation)
    if graph_inter:
        _set(vm.DMC, "forward_inter", _graphed_forward_inter(vm.DMC.forward_inter))
    if train_module is not None:
        _set(train_module, "collect_likelihoods_list", rate.collect_likelihoods_list)


def unpatch():
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)
