"""Generate tests/golden/coder.npz: the real-bitstream path of the UNMODIFIED
reference (``MotionContextModel.compress`` / ``decompress``,
/root/reference/dmc/models/video_model.py:236-291) executed on CPU in the build
container, with the oracle's CompressAI surface underneath (the real package is
absent: the bit-stream arithmetic is oracle/c/rans_ref.c -- "parity unpinned",
see that file's header; what these vectors pin is the reference's own call
sequence, tensor shapes, symbol/index order and table construction inputs).

    python tests/golden/make_golden_coder.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.load_reference import load_reference_models  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _bytes(s):
    return np.frombuffer(s, dtype=np.uint8).copy()


def main():
    torch.set_num_threads(1)
    load_reference_models()
    vm = sys.modules["models.video_model"]
    scale_table = sys.modules["models.base_model"].get_scale_table()
    torch.manual_seed(17)
    model = vm.MotionContextModel(ch_mv=8).eval()
    with torch.no_grad():        # spread the learned quantiles so the EB tables are ragged
        q = model.entropy_bottleneck.quantiles
        q[:, 0, 0] = -torch.linspace(2.5, 30, 8)
        q[:, 0, 2] = torch.linspace(1.5, 20, 8)
        q[:, 0, 1] = torch.linspace(-1, 1, 8)
    gc, eb = model.gaussian_conditional, model.entropy_bottleneck
    gc.update_scale_table(scale_table, force=True)
    eb.update(force=True)
    g = torch.Generator().manual_seed(18)
    y = torch.randn(2, 8, 16, 16, generator=g) * 6
    y[0, 0, 0, :4] += torch.tensor([900.0, -900.0, 70000.0, -70000.0])      # escapes
    y_ref = torch.randn(2, 8, 16, 16, generator=g)
    with torch.no_grad():
        y_hat, out = model.compress(y, y_ref)
        dec = model.decompress(out["strings"], out["shape"], y_ref)
        assert torch.equal(dec, y_hat)
        z = model.hyper_encoder(y)
        z_hat = eb.decompress(out["strings"][2], out["shape"])
        params = model.hyper_decoder(z_hat)
        means, scales = model.y_prior_fusion(torch.cat((params, y_ref), 1)).chunk(2, 1)
        _, q0, q1, s0, s1 = model.forward_dual_prior(y, means, scales, mode="compress")
        prior = model.y_spatial_prior(torch.cat(
            ((q0 + means.chunk(2, 1)[0]) * model.get_mask(16, 16, "cpu")[0],
             (q0 + means.chunk(2, 1)[1]) * model.get_mask(16, 16, "cpu")[1], means, scales), 1))
        i0, i1 = gc.build_indexes(s0), gc.build_indexes(s1)
    blob = {
        "scale_table": np.asarray(scale_table, dtype=np.float64),
        "gc.cdf": _np(gc._quantized_cdf), "gc.len": _np(gc._cdf_length), "gc.off": _np(gc._offset),
        "eb.cdf": _np(eb._quantized_cdf), "eb.len": _np(eb._cdf_length), "eb.off": _np(eb._offset),
        "y": _np(y), "z": _np(z), "z_hat": _np(z_hat), "means": _np(means), "scales": _np(scales),
        "prior": _np(prior), "q0": _np(q0), "q1": _np(q1), "s0": _np(s0), "s1": _np(s1),
        "i0": _np(i0), "i1": _np(i1), "y_hat": _np(y_hat),
        "shape": np.asarray(tuple(out["shape"]), dtype=np.int64),
    }
    for k, v in eb.state_dict().items():
        if v.numel() and k not in ("_offset", "_quantized_cdf", "_cdf_length"):
            blob[f"ebp.{k}"] = _np(v)
    for t, name in enumerate(("y0", "y1", "z")):
        for n, s in enumerate(out["strings"][t]):
            blob[f"str.{name}.{n}"] = _bytes(s)
    path = os.path.join(HERE, "coder.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), {k: len(v) for k, v in blob.items() if k.startswith("str.")})


if __name__ == "__main__":
    main()
