#!/bin/bash
# The sequence that produces what profiles/ holds (run on a B200 box through gpurun; TAG prefixes the output files):
#   gpurun --timeout 900 -- 'bash tools/refresh_profiles.sh r2'
# Plain runs first (bench values are never taken under a profiler), then the ncu launch list and one `--set full`
# capture of the kernels of one cfg4 step; condense afterwards with tools/ncu_summary.py.
TAG=${1:-v}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1 || { tail -20 $O/${TAG}_tests.log; echo 'GPU tests failed: not profiling'; exit 1; }
tail -1 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench_init.json 2> $O/${TAG}_bench.err
python bench.py --distribution trained --no-cpu-baseline > $O/${TAG}_bench_trained.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
for w in cfg3 cfg5 cfg2 cfg1; do python bench.py --workload $w --soak-seconds 0 > $O/${TAG}_bench_$w.json 2>> $O/${TAG}_bench.err; done
for w in cfg2 cfg1; do python bench.py --workload $w --cuda-graphs --no-cpu-baseline --soak-seconds 0 --steps 200 > $O/${TAG}_bench_${w}_graphs.json 2>> $O/${TAG}_bench.err; done
python tools/gpu_microbench.py --rows > $O/${TAG}_micro.jsonl 2>> $O/${TAG}_bench.err
python tools/fold_postconv_bench.py --workload cfg4 > $O/${TAG}_fold_cfg4.json 2>> $O/${TAG}_bench.err
for w in cfg4 cfg5; do python tools/fold_quantconv_bench.py --workload $w > $O/${TAG}_fold_quantconv_$w.json 2>> $O/${TAG}_bench.err; done
# (multi-GPU, separately: gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
#   --master-port 29533 tools/dp_overhead.py' -> profiles/r2_dp_overhead_8gpu.json; bench.py --gpus N under the same launcher)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --soak-seconds 0 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --soak-seconds 0 > $O/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vq_ --launch-skip 32 --launch-count 8 -f -o $O/${TAG}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --soak-seconds 0 > $O/${TAG}_ncu2.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>> $O/${TAG}_ncu2.log
tail -c 600 $O/${TAG}_bench_init.json
