"""The C ABI promises "launches only on the stream it is given, never synchronises,
CUDA-graph capturable" (include/dvc_b200.h).  Capture each entry
family in a CUDA graph, replay it on fresh inputs and compare with the eager launch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _capture(fn, warm=2):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def test_pframe_hot_path_in_a_cuda_graph(cuda_dev):
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200.pipeline import PFramePath, synthetic_pframe_inputs
    torch.manual_seed(1)
    ebs = {"motion": dvc.EntropyBottleneck(64).to(cuda_dev).eval(),
           "frame": dvc.EntropyBottleneck(64).to(cuda_dev).eval()}
    with torch.no_grad():
        inp = synthetic_pframe_inputs(128, 192, cuda_dev, seed=3)
        path = PFramePath(inp, ebs)
        eager = {k: v.clone() for k, v in path.launch(concurrent=False).items()
                 if isinstance(v, torch.Tensor)}
        torch.cuda.synchronize()
        g, out = _capture(lambda: path.launch(concurrent=True))
        # new data in the captured input buffers: the graph must recompute, not replay results
        fresh = synthetic_pframe_inputs(128, 192, cuda_dev, seed=4)
        for k, v in fresh.items():
            if isinstance(v, torch.Tensor):
                inp[k].copy_(v)
        g.replay()
        torch.cuda.synchronize()
        replayed = {k: v.clone() for k, v in out.items() if isinstance(v, torch.Tensor)}
        ref = PFramePath(fresh, ebs).launch(concurrent=False)
        torch.cuda.synchronize()
    assert any(not torch.equal(eager[k], replayed[k]) for k in ("context1", "bits"))
    for k, v in replayed.items():
        if k in ("mv2", "mv3"):      # intermediates of the reference function: never materialised
            continue
        assert torch.equal(v, ref[k]), k


def test_planar_warp_and_fused_conv_in_a_cuda_graph(cuda_dev):
    import deepvideocodec_b200 as dvc
    g0 = torch.Generator(device="cpu").manual_seed(5)
    feat = torch.randn(1, 64, 16, 136, generator=g0).to(cuda_dev)                 # NCHW: planar path
    feat_cl = feat.contiguous(memory_format=torch.channels_last)
    extra = torch.randn(1, 64, 16, 136, generator=g0).to(cuda_dev).contiguous(memory_format=torch.channels_last)
    flow = (torch.randn(1, 2, 16, 136, generator=g0) * 3).to(cuda_dev)
    weight = (torch.randn(64, 128, 3, 3, generator=g0) * 0.05).to(cuda_dev)
    bias = torch.randn(64, generator=g0).to(cuda_dev)
    with torch.no_grad():
        dvc.pack_conv3x3_weight(weight, 64)      # the cache fill allocates: do it outside the capture

        def step():
            return dvc.flow_warp(feat, flow), dvc.warp_conv3x3(feat_cl, flow, weight, bias, extra)
        g, (w_out, (ctx, conv)) = _capture(step)
        flow.copy_((torch.randn(1, 2, 16, 136, generator=g0) * 20).to(cuda_dev))   # incoherent now
        g.replay()
        torch.cuda.synchronize()
        ref_w = dvc.flow_warp(feat, flow)
        ref_ctx, ref_conv = dvc.warp_conv3x3(feat_cl, flow, weight, bias, extra)
    assert torch.equal(w_out, ref_w)
    assert torch.equal(ctx, ref_ctx) and torch.equal(conv, ref_conv)
