"""Model-driven golden vectors (SURVEY.md 8d "model-driven regime"): the stock,
UNMODIFIED reference ``DMC`` (random init, seed 0) codes three 64x64 frames on
the CPU and every tensor that crosses the hot-path boundary is recorded.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_model.py      ->  tests/golden/model_capture.npz

What is recorded, per P-frame (2 of them; the second one runs with a full
decoded-picture buffer, video_model.py:543-549):
* every ``flow_warp`` call of ``DMC.motion_compensation`` (video_model.py:497-506):
  input, flow, output (64-channel features: channels 0..7 only, the warp is
  per-channel) and both ``bilineardownsacling`` calls;
* for both context models: ``forward_dual_prior`` inputs and outputs
  (video_model.py:169-216 / 341-388), the spatial-prior conv output in
  between, the Gaussian-conditional call (inputs, scales, means -> likelihood),
  the entropy-bottleneck call (z -> outputs, likelihood) and the hyper-latent
  ``z_hat`` (:222-224 / 394-396);
* the entropy-bottleneck parameters, so the test can rebuild the module;
* ``collect_likelihoods_list`` of the whole output (train.py:74-93).

``compressai`` is the oracle shim (the real package is absent), so the two
likelihood arithmetics here are the restatement's -- "parity unpinned", see
oracle/__init__.py; everything else is the reference's own code and ATen.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.load_reference import load_reference_models, load_reference_train_fn  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy().copy()


def main():
    torch.set_num_threads(1)
    load_reference_models()
    vm = sys.modules["models.video_model"]
    layers = sys.modules["models.layers"]
    collect = load_reference_train_fn("collect_likelihoods_list")
    rec = {}
    state = {"frame": 0, "warp": 0, "down": 0}

    # ---- wrap the names DMC.motion_compensation resolves at call time ------------
    ref_warp, ref_down = vm.flow_warp, vm.bilineardownsacling

    def warp(im, flow):
        layers.backward_grid[-1].clear()        # the reference cache is keyed by shape only
        out = ref_warp(im, flow)
        k = f"f{state['frame']}.warp{state['warp']}"
        c = min(im.size(1), 8)
        rec[k + ".im"], rec[k + ".flow"], rec[k + ".out"] = _np(im[:, :c]), _np(flow), _np(out[:, :c])
        state["warp"] += 1
        return out

    def down(x):
        out = ref_down(x)
        k = f"f{state['frame']}.down{state['down']}"
        rec[k + ".in"], rec[k + ".out"] = _np(x), _np(out)
        state["down"] += 1
        return out

    vm.flow_warp, vm.bilineardownsacling = warp, down

    # ---- context models ----------------------------------------------------------------
    def wrap_context(cls, label):
        ref_fdp = cls.forward_dual_prior
        ref_fwd = cls.forward

        def fdp(self, y, means, scales, mode="trainval"):
            out = ref_fdp(self, y, means, scales, mode)
            k = f"f{state['frame']}.{label}"
            rec[k + ".y"], rec[k + ".means"], rec[k + ".scales"] = _np(y), _np(means), _np(scales)
            rec[k + ".y_hat"], rec[k + ".means_hat"], rec[k + ".scales_hat"] = (_np(t) for t in out[:3])
            return out

        def fwd(self, *args):
            k = f"f{state['frame']}.{label}"
            def on_prior(m, i, o):
                rec[k + ".prior"] = _np(o)

            def on_gc(m, i, o):
                rec[k + ".y_lik"] = _np(o[1])

            def on_eb(m, i, o):
                rec[k + ".z"], rec[k + ".z_out"], rec[k + ".z_lik"] = _np(i[0]), _np(o[0]), _np(o[1])

            def on_hyper_decoder(m, i, o):
                rec[k + ".z_hat"] = _np(i[0])
            hooks = [self.y_spatial_prior.register_forward_hook(on_prior),
                     self.gaussian_conditional.register_forward_hook(on_gc),
                     self.entropy_bottleneck.register_forward_hook(on_eb),
                     self.hyper_decoder.register_forward_hook(on_hyper_decoder)]
            try:
                return ref_fwd(self, *args)
            finally:
                for h in hooks:
                    h.remove()
        cls.forward_dual_prior, cls.forward = fdp, fwd

    wrap_context(vm.MotionContextModel, "motion")
    wrap_context(vm.FrameContextModel, "frame")

    ref_inter = vm.DMC.forward_inter

    def inter(self, *a, **kw):
        state["warp"] = state["down"] = 0
        out = ref_inter(self, *a, **kw)
        state["frame"] += 1
        return out
    vm.DMC.forward_inter = inter

    torch.manual_seed(0)
    net = vm.DMC().eval()
    g = torch.Generator().manual_seed(20261018)
    frames = [torch.rand(1, 3, 64, 64, generator=g) for _ in range(3)]
    with torch.no_grad():
        out = net(frames)
    assert state["frame"] == 2
    for label, cm in (("motion", net.motion_context_model), ("frame", net.frame_context_model)):
        for name, p in cm.entropy_bottleneck.state_dict().items():
            rec[f"eb.{label}.{name}"] = _np(p)
    num_pixels = 64 * 64 * 2                     # train.py:172: H * W * number of P-frames
    bpp, info = collect(out["likelihoods"], num_pixels)
    rec["num_pixels"] = np.int64(num_pixels)
    rec["bpp_loss"] = _np(bpp)
    for k, v in info.items():
        rec["info." + k] = np.float64(float(v))
    floor = np.mean([np.mean(rec[f"f{f}.{m}.y_lik"] <= 1e-9) for f in range(2) for m in ("motion", "frame")])
    print(f"{len(rec)} arrays, {sum(v.nbytes for v in rec.values()) / 1e6:.2f} MB raw, "
          f"fraction of y likelihoods on the 1e-9 floor: {floor:.2f}, bpp {float(bpp):.3f}")
    np.savez_compressed(os.path.join(HERE, "model_capture.npz"), **rec)


if __name__ == "__main__":
    main()
