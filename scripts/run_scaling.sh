set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/s1.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_line_${N}gpu.json 2> gpurun_out/s$N.err
done
for W in config2 config4 config1; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r2_bench_line_${W}.json 2> gpurun_out/w1_$W.err
  $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --workload $W --steps 10 --warmup 3 > gpurun_out/r2_bench_line_${W}_8gpu.json 2> gpurun_out/w8_$W.err
done
python scripts/sweep.py > gpurun_out/r2_sweep_1gpu.jsonl 2> gpurun_out/sw1.err
$TR --nproc-per-node 8 --master-port 29530 scripts/sweep.py > gpurun_out/r2_sweep_8gpu.jsonl 2> gpurun_out/sw8.err
python -m pytest tests/test_multidevice_gpu.py tests/test_siteblocks_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_multigpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_l.log 2>&1
tail -3 gpurun_out/r2_multigpu_tests.log
