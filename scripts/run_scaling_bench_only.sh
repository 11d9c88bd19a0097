# config 3 at 1/2/4/8 GPUs on one box (bench lines only)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs | grep '^{' > gpurun_out/r2_bench_line_1gpu.json
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_bench_line_${N}gpu.json
done
$TR --nproc-per-node 8 --master-port 29620 bench.py --gpus 8 --workload config2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_bench_line_config2_8gpu.json
$TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/r2_bench_line_reference_arm_under_torchrun8.json
python - <<'PY'
import json
for f in ("1gpu","2gpu","4gpu","8gpu","config2_8gpu"):
    d=json.load(open(f"gpurun_out/r2_bench_line_{f}.json"))
    print(f, d["n_gpus"], round(d["ms_per_step"],4), round(d["value"],2), round(d["e2e"]["value"],2), d["elbo"], d["grad_checksum"], round(d["roofline"]["kernel_share_of_step"],4), d["clocks"]["sm_mhz"] if d["clocks"] else None)
d=json.load(open("gpurun_out/r2_bench_line_reference_arm_under_torchrun8.json")); print("ref@torchrun8", d["value"], d["cpu_baseline"]["cores"], d["native_so_loaded"])
PY
