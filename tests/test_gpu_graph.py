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


def test_whole_pframe_of_the_reference_under_one_graph(cuda_dev):
    """SURVEY.md 8f row f4 as a product feature: the reference's own ``DMC.forward_inter``
    (patched) replayed from ONE CUDA graph per P-frame (``dvc.GraphedInter``, what
    ``patch(models, graph_inter=True)`` binds) is bit-identical to the eager patched call over
    a 4-frame GOP (first P-frame with an empty dpb, then populated ones), follows in-place
    parameter updates (entropy-bottleneck packing is a host-side cache) and falls back to
    eager under autograd."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import dropin_util as du
    import deepvideocodec_b200 as dvc
    if not du.reference_available():
        pytest.skip("reference sources not staged (python tools/stage_reference.py) -- "
                    "THE GRAPHED P-FRAME TEST DID NOT RUN")
    with du.deterministic_convs():
        _, model = du.build_pair(cuda_dev, seed=0, weight_scale=0.7)
        model.eval()
        fr = du.frames(4, 1, 128, 192, cuda_dev, seed=9)
        graphed = dvc.GraphedInter(model)

        def run(inter):
            dpb = {"x_ref": fr[0], "feature_ref": None, "y_ref": None, "y_mv_ref": None}
            outs = []
            for x in fr[1:]:
                x_rec, lik, ctx = inter(x, dpb)
                outs.append((x_rec, lik))
                dpb = {"x_ref": x_rec, "feature_ref": ctx["feature_ref"], "y_ref": ctx["y_ref"],
                       "y_mv_ref": ctx["y_mv_ref"]}
            return outs

        def same(a, b):
            for (xa, la), (xb, lb) in zip(a, b):
                assert torch.equal(xa, xb)
                for label in ("motion", "frame"):
                    for f in ("y", "z"):
                        assert torch.equal(la[label][f], lb[label][f])
                        assert torch.equal(la[label][f]._dvc_logsum, lb[label][f]._dvc_logsum)

        with torch.no_grad():
            eager = run(lambda x, d: model.forward_inter(x, d))
            g1 = run(graphed)
            g2 = run(graphed)                      # replays only
            assert len(graphed._graphs) == 2       # empty dpb, populated dpb
            same(eager, g1)
            same(eager, g2)
            # results are clones: frame 1's x_rec survives the later replays
            assert torch.equal(g1[0][0], eager[0][0])
            # an in-place update of an entropy-bottleneck parameter must show up
            eb = model.frame_context_model.entropy_bottleneck
            eb._bias0.add_(0.25)
            try:
                same(run(lambda x, d: model.forward_inter(x, d)), run(graphed))
                assert not torch.equal(run(graphed)[0][1]["frame"]["z"], eager[0][1]["frame"]["z"])
            finally:
                eb._bias0.sub_(0.25)
        # autograd on -> the eager method (a graph cannot record a backward)
        x_rec, _, _ = graphed(fr[1], {"x_ref": fr[0], "feature_ref": None, "y_ref": None,
                                      "y_mv_ref": None})
        assert x_rec.requires_grad
