set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_gop.py -m gpu -q > gpurun_out/r2_gop_t8.log 2>&1; tail -3 gpurun_out/r2_gop_t8.log
for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2951$N tools/h2d_probe.py > gpurun_out/r2_h2d_N$N.log 2>&1; done
python bench.py --gpus 1 --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_scale_N1.json 2> gpurun_out/r2_scale_N1.err
for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/r2_scale_N$N.json 2> gpurun_out/r2_scale_N$N.err; done
python bench.py --gop --sequences 2 > gpurun_out/r2_gop_N1.json 2> gpurun_out/r2_gop_N1.err
for N in 2 8; do $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --gop --sequences 2 > gpurun_out/r2_gop_N$N.json 2> gpurun_out/r2_gop_N$N.err; done
for N in 2 8; do $TR --nproc-per-node $N --master-port 2954$N tools/train_ddp_bench.py > gpurun_out/r2_ddp_N$N.log 2>&1; done
for N in 1 2 4 8; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2_scale_N$N.json'))
print($N, round(d['value'],1), round(d['e2e']['value'],1), d['e2e']['copy_only'], d['e2e'].get('all_inputs_from_host'))
"; done
cat gpurun_out/r2_gop_N*.json | cut -c1-400
tail -1 gpurun_out/r2_ddp_N8.log | cut -c1-2500
