#!/bin/bash
# Multi-GPU measurements of one round (run under `gpurun --gpus N -- bash tools/run_multi.sh N TAG`): the default bench line
# (config 2), the sharded inflate (config 4, with and without the gather), configs 3 and 5, and the multi-GPU parity test.
N=$1; TAG=${2:-r02}
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@"; }
run 29611 --steps 3 --warmup 3                                   > gpurun_out/${TAG}_bench_config2_n$N.json 2> gpurun_out/${TAG}_bench_config2_n$N.err
run 29612 --config 4 --steps 3 --warmup 3 --no-cpu               > gpurun_out/${TAG}_bench_config4_n$N.json 2> gpurun_out/${TAG}_bench_config4_n$N.err
run 29613 --config 4 --steps 3 --warmup 3 --no-cpu --gather      > gpurun_out/${TAG}_bench_config4_gather_n$N.json 2> gpurun_out/${TAG}_bench_config4_gather_n$N.err
run 29614 --config 3 --steps 2 --warmup 3 --no-cpu               > gpurun_out/${TAG}_bench_config3_n$N.json 2> gpurun_out/${TAG}_bench_config3_n$N.err
run 29615 --config 5 --steps 2 --warmup 3 --no-cpu --no-e2e      > gpurun_out/${TAG}_bench_config5_n$N.json 2> gpurun_out/${TAG}_bench_config5_n$N.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/${TAG}_multi_gpu_parity_n$N.txt 2>&1
for f in gpurun_out/${TAG}_bench_config*_n$N.json; do echo "$f: $(cut -c1-150 $f)"; done
tail -3 gpurun_out/${TAG}_multi_gpu_parity_n$N.txt
