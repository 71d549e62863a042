"""One CUDA graph per P-frame.

The reference's ``DMC.forward_inter`` (``dmc/models/video_model.py:556-579``) is
~1 450 eager kernel launches per P-frame (181 convolutions + the hot path's
~300 element-wise launches + a host->device mask upload per context model,
``:154``).  With ``deepvideocodec_b200.patch`` applied the non-conv part is a
dozen launches with no host synchronisation, no host tensor and no allocation
outside the caching allocator -- so the whole call captures into ONE CUDA graph
and replays with a single launch.  Measured on the reference's own ``DMC`` on a
B200 (``profiles/r02_dropin.json``): 256x256 (BASELINE.json configs[0]) 14.4 ms
stock eager -> 11.1 ms patched -> 4.97 ms graphed; 1088x1920 71.1 -> 67.4 ->
65.2 ms (the 1080p frame is fp32 convolution time).

``GraphedInter(model)`` is a drop-in callable for ``model.forward_inter`` in
inference (``torch.no_grad()``, ``model.eval()``); ``patch(models,
graph_inter=True)`` binds it onto ``DMC.forward_inter`` so that the reference's
``DMC.forward`` / ``test_epoch`` (``train.py:349-397``) use it unchanged.
Outside those conditions (autograd on, training mode, CPU tensors) the original
method runs.  Results are bit-identical to the eager patched call
(``tests/test_gpu_graph.py``).
"""
import torch

__all__ = ["GraphedInter"]

_DPB_KEYS = ("x_ref", "feature_ref", "y_ref", "y_mv_ref")


def _sig(t):
    return None if t is None else (tuple(t.shape), tuple(t.stride()), t.dtype, t.device.index)


class _Captured:
    __slots__ = ("graph", "x_cur", "dpb", "out")


class GraphedInter:
    """``GraphedInter(model)(x_cur, dpb, motion_pretrain=False, frame_pretrain=False)``
    == ``model.forward_inter(...)`` with the whole call replayed from a CUDA
    graph.  One graph per (input shapes / strides, which dpb entries are None,
    flags) signature; static input buffers are filled with ``copy_`` before a
    replay and every returned tensor is a fresh clone, so callers may keep
    results across frames exactly as with the eager call (``DMC.forward``
    appends ``x_rec`` and the likelihood dicts to lists, ``:541-542``)."""

    def __init__(self, model, eager_inter=None, warmup=2):
        self.model = model
        self._eager = eager_inter if eager_inter is not None else type(model).forward_inter
        self._warmup = int(warmup)
        self._graphs = {}
        # Host-side caches that a replay cannot refresh: the packed entropy-bottleneck
        # parameters (entropy_models._packed_cached) are rebuilt by Python when a parameter's
        # version changes.  Their (address, version) fingerprint is part of the graph key, so
        # an optimiser step or load_state_dict between two evaluations re-captures.  Conv
        # weights are read in place by the replayed kernels and need nothing.
        self._watched = [p for m in model.modules() if hasattr(m, "_matrix0")
                         for p in m.parameters(recurse=False)]
        first = next(iter(model.parameters()), None)
        if first is not None:
            self._watched.append(first)           # catches a wholesale .to() / re-allocation

    def _fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self._watched)

    # -- helpers ---------------------------------------------------------------------
    @staticmethod
    def _clone_tree(obj):
        if isinstance(obj, torch.Tensor):
            c = obj.clone()
            ls = getattr(obj, "_dvc_logsum", None)
            if ls is not None:                    # fused sum(ln p) travels with the likelihood
                c._dvc_logsum = ls.clone()
            return c
        if isinstance(obj, dict):
            return {k: GraphedInter._clone_tree(v) for k, v in obj.items()}
        if isinstance(obj, (tuple, list)):
            return type(obj)(GraphedInter._clone_tree(v) for v in obj)
        return obj

    def _capture(self, x_cur, dpb, flags):
        cap = _Captured()
        cap.x_cur = x_cur.clone()
        cap.dpb = {k: (None if dpb.get(k) is None else dpb[k].clone()) for k in _DPB_KEYS}
        side = torch.cuda.Stream(device=x_cur.device)
        side.wait_stream(torch.cuda.current_stream(x_cur.device))
        with torch.cuda.stream(side):             # warm-up: caches, cuDNN plans, workspaces
            for _ in range(self._warmup):
                self._eager(self.model, cap.x_cur, cap.dpb, *flags)
        torch.cuda.current_stream(x_cur.device).wait_stream(side)
        torch.cuda.synchronize(x_cur.device)
        cap.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cap.graph):
            cap.out = self._eager(self.model, cap.x_cur, cap.dpb, *flags)
        return cap

    # -- the call -----------------------------------------------------------------------
    def __call__(self, x_cur, dpb, motion_pretrain=False, frame_pretrain=False):
        flags = (bool(motion_pretrain), bool(frame_pretrain))
        if (torch.is_grad_enabled() or self.model.training or not x_cur.is_cuda
                or torch.cuda.is_current_stream_capturing()):
            return self._eager(self.model, x_cur, dpb, *flags)
        key = (_sig(x_cur), tuple(_sig(dpb.get(k)) for k in _DPB_KEYS), flags, self._fingerprint())
        cap = self._graphs.get(key)
        if cap is None:
            if len(self._graphs) >= 8:            # stale weight versions: do not hoard graph pools
                self._graphs.clear()
            cap = self._graphs[key] = self._capture(x_cur, dpb, flags)
        cap.x_cur.copy_(x_cur)
        for k in _DPB_KEYS:
            if cap.dpb[k] is not None:
                cap.dpb[k].copy_(dpb[k])
        cap.graph.replay()
        return self._clone_tree(cap.out)

    def clear(self):
        """Drop every captured graph (e.g. after loading new weights into
        *reallocated* parameters; in-place ``load_state_dict`` needs nothing)."""
        self._graphs.clear()
