#!/bin/bash
# One 8 x B200 box (gpurun --gpus 8): the scaling curve of bench.py (device-resident value, e2e,
# copy-only ceiling), the H2D probe, BASELINE configs[2] (--gop) and configs[3] (DDP training).
# Outputs land in gpurun_out/; profiles/r02_{scale,h2d_probe,gop_scale,train_ddp}.json are built
# from them.  PROBES=0 skips the probe and the DDP bench (already recorded).
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_gop.py -m gpu -q > gpurun_out/r2_gop_t8.log 2>&1; tail -3 gpurun_out/r2_gop_t8.log
if [ "${PROBES:-1}" = "1" ]; then
  for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2951$N tools/h2d_probe.py > gpurun_out/r2_h2d_N$N.log 2>&1; done
  for N in 2 8; do $TR --nproc-per-node $N --master-port 2954$N tools/train_ddp_bench.py > gpurun_out/r2_ddp_N$N.log 2>&1; done
fi
python bench.py --gpus 1 --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_scale_N1.json 2> gpurun_out/r2_scale_N1.err
for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/r2_scale_N$N.json 2> gpurun_out/r2_scale_N$N.err; done
python bench.py --gop --sequences 2 > gpurun_out/r2_gop_N1.json 2> gpurun_out/r2_gop_N1.err
for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --gop --sequences 2 > gpurun_out/r2_gop_N$N.json 2> gpurun_out/r2_gop_N$N.err; done
grep -h -o '"value": [0-9.]*' gpurun_out/r2_scale_N*.json gpurun_out/r2_gop_N*.json
