TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_scale3_N1.json 2> gpurun_out/r2_scale3_N1.err
for N in 2 4 8; do $TR --nproc-per-node $N --master-port 2962$N bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/r2_scale3_N$N.json 2> gpurun_out/r2_scale3_N$N.err; done
$TR --nproc-per-node 8 --master-port 29639 bench.py --gpus 8 --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
grep -h -o '"value": [0-9.]*' gpurun_out/r2_scale3_N*.json | head -20
