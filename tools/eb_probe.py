#!/usr/bin/env python
"""Which accumulation order does eager ``torch.matmul`` (cuBLAS batched GEMM, K <= 3) use in
the entropy bottleneck's logits (CompressAI ``_logits_cumulative``) at a given
column count?  The kernel replays an ascending-k FMA chain; at 1080p (510
columns) that is bit-identical to eager, at 256x256 (16 columns) a few tail
likelihoods differ by ~1e-5 relative (tests/test_gpu_dropin.py).  This probe
evaluates candidate orders in fp64-emulated fp32 and counts mismatches against
``torch.matmul`` per layer and column count.  Output: gpurun_out/eb_probe.json
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def r32(x):
    return x.to(torch.float32).to(torch.float64)


def fma(a, b, c):
    """round32(a*b + c) with a, b, c fp32 values held in fp64 (a*b exact)."""
    return r32(a * b + c)


def candidates(A, B):
    """A [C,M,K], B [C,K,N] fp32 -> dict name -> [C,M,N] fp32."""
    a, b = A.double(), B.double()
    K = A.size(2)
    out = {}
    prods = [a[:, :, k:k + 1] * b[:, k:k + 1, :] for k in range(K)]        # exact in fp64
    acc = r32(prods[0])
    for k in range(1, K):
        acc = fma(a[:, :, k:k + 1], b[:, k:k + 1, :], acc)
    out["fma_ascending"] = acc
    acc = r32(prods[K - 1])
    for k in range(K - 2, -1, -1):
        acc = fma(a[:, :, k:k + 1], b[:, k:k + 1, :], acc)
    out["fma_descending"] = acc
    acc = r32(prods[0])
    for k in range(1, K):
        acc = r32(acc + r32(prods[k]))
    out["separate_ascending"] = acc
    if K == 3:
        out["fma_pair_02_1"] = fma(a[:, :, 1:2], b[:, 1:2, :],
                                   fma(a[:, :, 2:3], b[:, 2:3, :], r32(prods[0])))
        out["exact_sum"] = r32(prods[0] + prods[1] + prods[2])
    if K == 3 and A.size(1) == 1:
        import itertools
        P = [r32(p) for p in prods]
        aa = [a[:, :, k:k + 1] for k in range(3)]
        bb = [b[:, k:k + 1, :] for k in range(3)]
        for i, j, k in itertools.permutations(range(3)):
            out[f"fma{k}(fma{j}(P{i}))"] = fma(aa[k], bb[k], fma(aa[j], bb[j], P[i]))
            out[f"fma{j}(P{i})+P{k}"] = r32(fma(aa[j], bb[j], P[i]) + P[k])
            if i < j:
                out[f"(P{i}+P{j})+P{k}"] = r32(r32(P[i] + P[j]) + P[k])
                out[f"fma{k}(P{i}+P{j})"] = fma(aa[k], bb[k], r32(P[i] + P[j]))
                out[f"fma{k}(fma{j}(P{i}))+0"] = out[f"fma{k}(fma{j}(P{i}))"]
    return {k: v.float() for k, v in out.items()}


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    res = {}
    C = 64
    # where does the [C,1,3] @ [C,3,N] product change kernels?
    switch = {}
    for C2 in (64, 128, 192):
        last_bad = None
        for ncols in list(range(100, 600, 10)) + [768, 1024]:
            A = F.softplus(torch.randn(C2, 1, 3, device=dev))
            B = torch.randn(C2, 3, ncols, device=dev) * 10
            bad = int((candidates(A, B)["fma_ascending"] != torch.matmul(A, B)).sum().item())
            if bad:
                last_bad = ncols
        switch[str(C2)] = last_bad
    print("last column count with a non-ascending-FMA 1x3 product, per batch:", switch, flush=True)
    res["switch_1x3"] = switch
    for ncols in (16, 120, 510):
        row = {}
        for (m, k) in ((3, 1), (3, 3), (1, 3)):
            A = F.softplus(torch.randn(C, m, k, device=dev))
            B = torch.randn(C, k, ncols, device=dev) * 10
            ref = torch.matmul(A, B)
            cand = candidates(A, B)
            row[f"{m}x{k}"] = {n: int((v != ref).sum().item()) for n, v in cand.items()}
            row[f"{m}x{k}"]["elements"] = ref.numel()
        res[str(ncols)] = row
        print(ncols, json.dumps(row), flush=True)
    # the failing regime itself: which field, which values
    import dropin_util as du
    if du.reference_available():
        with du.deterministic_convs(), torch.no_grad():
            stock, patched = du.build_pair(dev, seed=0, weight_scale=1.0)
            stock.eval(), patched.eval()
            fr = du.frames(3, 1, 256, 256, dev, seed=1)
            zs = []
            hook = stock.motion_context_model.entropy_bottleneck.register_forward_hook(
                lambda m, i, o: zs.append(i[0].detach()))
            out_s, _ = du.run_forward(stock, fr)
            hook.remove()
            out_p, _ = du.run_forward(patched, fr)
            detail = []
            for i in range(2):
                for label in ("motion", "frame"):
                    for field in ("y", "z"):
                        a = out_p["likelihoods"][i][label][field].double()
                        b = out_s["likelihoods"][i][label][field].double()
                        rel = ((a - b).abs() / b)
                        idx = int(rel.argmax())
                        detail.append({"frame": i, "label": label, "field": field,
                                       "max_rel": float(rel.max()), "n_over_1e-5": int((rel > 1e-5).sum()),
                                       "n_differ": int((a != b).sum()), "numel": a.numel(),
                                       "stock_at_max": float(b.reshape(-1)[idx]),
                                       "patched_at_max": float(a.reshape(-1)[idx])})
            res["dropin_256_stock_init"] = detail
            res["z_abs_max"] = [float(z.abs().max()) for z in zs]
            for d in detail:
                print(json.dumps(d), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "eb_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
