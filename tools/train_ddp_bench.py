#!/usr/bin/env python
"""BASELINE.json configs[3]: DMC training step on 256x256 crops, 7-frame GOP,
batch 8 per GPU, noise-quantisation likelihoods, N x B200 DistributedDataParallel
(VERDICT r1 row g2).  Replaces the reference's inert ``nn.DataParallel``
(train.py:598-600; CUDA_VISIBLE_DEVICES is hard-coded to one GPU, train.py:43).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29521 tools/train_ddp_bench.py

Both arms run the reference's OWN ``DMC`` (staged by tools/stage_reference.py),
its own ``RateDistortionLoss`` / ``configure_optimizers`` / ``compute_aux_loss``
(train.py:96-211, 240-282) and one step exactly as ``train_one_epoch`` does
(train.py:285-337, the all-modules stage of :318-330):

  stock    unmodified reference over eager PyTorch ops
  patched  dvc.patch(models, train_module): warps, dual prior, likelihoods, rate and their
           backward passes are this package's kernels

DDP notes: ``*.quantiles`` are excluded from gradient reduction --
the aux loss (train.py:336) depends only on parameters, which are identical on
every rank, so its gradient needs no all-reduce; stages with unused sub-nets
(motion/frame pre-training, train.py:298-316) run with find_unused_parameters.

Reports, per arm: step ms (max over ranks, CUDA events), loss, gradient
all-reduce bytes, NCCL kernel time per step, the share of the step spent in the
hot-path ops, and parity: patched loss vs stock loss on the same batch / noise
(1e-4 rel), DDP-averaged gradient vs the mean of the per-rank local gradients.
Output (rank 0): gpurun_out/train_ddp_N<world>.json
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--weight-scale", type=float, default=0.7)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29521")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import deepvideocodec_b200 as dvc
    import dropin_util as du
    from oracle.load_reference import load_reference_train_ns
    from torch.nn.parallel import DistributedDataParallel as DDP
    dvc.lib()
    names = ["collect_likelihoods_list", "RateDistortionLoss", "compute_aux_loss",
             "configure_optimizers"]
    train_stock = load_reference_train_ns(names)
    train_patched = load_reference_train_ns(names)
    stock, patched = du.build_pair(dev, seed=0, weight_scale=args.weight_scale)
    # what dvc.patch(models, train_module=train) does for the reference's train.py
    train_patched.collect_likelihoods_list = dvc.collect_likelihoods_list

    class A:                                    # argparse stand-in for configure_optimizers
        learning_rate = 1e-4
        aux_learning_rate = 1e-3

    fr = du.frames(args.frames, args.batch, args.size, args.size, dev, seed=100 + rank)
    n_param = sum(p.numel() for n, p in stock.named_parameters() if not n.endswith(".quantiles"))
    res = {"world": world, "batch_per_gpu": args.batch, "frames": args.frames,
           "crop": [args.size, args.size], "params_reduced": n_param,
           "allreduce_bytes_per_step": 4 * n_param, "arms": {}}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def one_step(model, ddp, crit, train_ns, opt, aux_opt, seed, stage="all"):
        opt.zero_grad(set_to_none=True)
        aux_opt.zero_grad(set_to_none=True)
        torch.manual_seed(seed)                               # training noise: same in both arms
        kw = {"motion_pretrain": stage == "motion", "frame_pretrain": stage == "frame"}
        # motion pre-training returns no dpb context (video_model.py:565-566), so the reference
        # itself can only run it on 2-frame samples (:543-549 would raise KeyError)
        seq = fr[:2] if stage == "motion" else fr
        out = ddp(list(seq), **kw)
        oc = crit(out, seq[1:])
        oc["loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)      # train.py:332-333
        opt.step()
        aux = train_ns.compute_aux_loss(model.aux_loss(), backward=True)
        aux_opt.step()
        return oc["loss"].detach(), aux.detach()

    for name, model, train_ns in (("stock", stock, train_stock), ("patched", patched, train_patched)):
        model.train()
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        model._ddp_params_and_buffers_to_ignore = [
            n for n, _ in model.named_parameters() if n.endswith(".quantiles")]
        ddp = DDP(model, device_ids=[local], broadcast_buffers=False,
                  gradient_as_bucket_view=True) if world > 1 else model
        crit = train_ns.RateDistortionLoss(lmbda=1e-2, return_details=True)
        opt, aux_opt = train_ns.configure_optimizers(model, A)
        loss0, aux0 = one_step(model, ddp, crit, train_ns, opt, aux_opt, seed=7)   # first step: parity
        for i in range(args.warmup - 1):
            one_step(model, ddp, crit, train_ns, opt, aux_opt, seed=8 + i)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for i in range(args.steps):
            loss, aux = one_step(model, ddp, crit, train_ns, opt, aux_opt, seed=100 + i)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        arm = {"step_ms": float(t), "first_step_loss": float(loss0), "first_step_aux": float(aux0),
               "last_loss": float(loss),
               "samples_per_s": world * args.batch / (float(t) * 1e-3),
               "p_frames_per_s": world * args.batch * (args.frames - 1) / (float(t) * 1e-3)}
        # kernel-time breakdown of one step (rank 0)
        if rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                one_step(model, ddp, crit, train_ns, opt, aux_opt, seed=999)
                torch.cuda.synchronize()
            tot = own = nccl = 0.0
            n_k = 0
            top = {}
            for evn in prof.events():
                if evn.device_type != torch.autograd.DeviceType.CUDA:
                    continue
                us = evn.device_time
                nm = evn.name
                if "Memcpy" in nm or "Memset" in nm:
                    continue
                n_k += 1
                tot += us
                if "nccl" in nm.lower():
                    nccl += us
                if "dvc::" in nm or "wc::" in nm:
                    own += us
                    top[nm.split("(")[0][:60]] = top.get(nm.split("(")[0][:60], 0.0) + us
            arm.update({"kernel_launches": n_k, "kernel_time_ms": tot / 1e3,
                        "nccl_kernel_ms": nccl / 1e3, "own_kernel_ms": own / 1e3,
                        "own_kernels_top": dict(sorted(top.items(), key=lambda kv: -kv[1])[:8])})
        else:
            one_step(model, ddp, crit, train_ns, opt, aux_opt, seed=999)
        # DDP gradient == mean over ranks of the local gradients (one parameter, fresh weights)
        if world > 1:
            model.load_state_dict(state0)
            probe = next(p for n, p in model.named_parameters()
                         if n.startswith("motion_decoder") and p.dim() == 4)
            with ddp.no_sync():
                model.zero_grad(set_to_none=True)
                torch.manual_seed(7)
                out = ddp(list(fr))
                crit(out, fr[1:])["loss"].backward()
            local_g = probe.grad.detach().clone()
            mean_g = local_g.clone()
            dist.all_reduce(mean_g, op=dist.ReduceOp.SUM)
            mean_g /= world
            model.zero_grad(set_to_none=True)
            torch.manual_seed(7)
            out = ddp(list(fr))
            crit(out, fr[1:])["loss"].backward()
            err = (probe.grad - mean_g).abs().max() / mean_g.abs().max().clamp_min(1e-30)
            arm["ddp_grad_vs_mean_of_local_rel"] = float(err)
        del ddp, opt, aux_opt
        res["arms"][name] = arm
        torch.cuda.empty_cache()

    s, p = res["arms"]["stock"], res["arms"]["patched"]
    res["loss_rel_err_patched_vs_stock_first_step"] = abs(p["first_step_loss"] - s["first_step_loss"]) / abs(s["first_step_loss"])
    res["aux_equal"] = p["first_step_aux"] == s["first_step_aux"]
    res["speedup_step"] = s["step_ms"] / p["step_ms"]
    if rank == 0 and "kernel_time_ms" in s:
        # hot-path share: stock kernel time not accounted for by the (identical) conv / optimizer
        # work of the patched arm = time of the eager hot-path ops (derived, not attributed by name)
        rest = p["kernel_time_ms"] - p["own_kernel_ms"] - p["nccl_kernel_ms"]
        res["hot_path_share_stock"] = max(0.0, s["kernel_time_ms"] - s["nccl_kernel_ms"] - rest) / \
            max(1e-9, s["kernel_time_ms"] - s["nccl_kernel_ms"])
        res["hot_path_share_patched"] = p["own_kernel_ms"] / max(1e-9, p["kernel_time_ms"] - p["nccl_kernel_ms"])
    # pre-training stages leave sub-nets unused (train.py:298-316): they need find_unused_parameters
    if world > 1:
        try:
            model = patched
            model.train()
            ddp = DDP(model, device_ids=[local], broadcast_buffers=False,
                      find_unused_parameters=True)
            crit = train_patched.RateDistortionLoss(lmbda=1e-2)
            opt, aux_opt = train_patched.configure_optimizers(model, A)
            for stage in ("motion", "frame"):
                loss, _ = one_step(model, ddp, crit, train_patched, opt, aux_opt, 5, stage=stage)
                res[f"stage_{stage}_pretrain_loss"] = float(loss)
        except Exception as e:  # noqa: BLE001
            res["stage_pretrain_error"] = repr(e)[:300]
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"train_ddp_N{world}.json"), "w") as fh:
            json.dump(res, fh, indent=1)
        print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
