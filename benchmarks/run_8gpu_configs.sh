# one 8-GPU lease: BASELINE configs[2] / configs[3] parity against the committed oracle vectors, then their bench lines
mkdir -p gpurun_out
T=tests/test_multigpu.py
timeout 600 python -m pytest "$T::test_baseline_config_vs_oracle_world" "$T::test_tiny_train_steps_vs_reference_world[2-2-fp32]" "$T::test_tiny_train_steps_vs_reference_world[2-2-bf16]" "$T::test_c1_shape_vs_oracle_world[2-2-bf16-1]" "$T::test_c1_shape_vs_oracle_world[2-2-fp32-1]" -m gpu -v --timeout=280 --timeout-method=thread > gpurun_out/r2_mg8.log 2>&1; echo "mg8 rc=$?"; grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/r2_mg8.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 bench.py --gpus 8 --config c3 --sweep --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n8.json 2> gpurun_out/r2_bench_c3_n8.err; echo "c3 rc=$?"; cut -c1-250 gpurun_out/r2_bench_c3_n8.json
timeout 240 $TR --master-port 29522 bench.py --gpus 8 --config c4 --sweep --steps 6 --warmup 3 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err; echo "c4 rc=$?"; cut -c1-250 gpurun_out/r2_bench_c4_n8.json
timeout 150 $TR --master-port 29523 bench.py --gpus 8 --config c4 --dp-sync all --steps 6 --warmup 3 --no-parity > gpurun_out/r2_bench_c4_n8_dpall.json 2> gpurun_out/r2_bench_c4_n8_dpall.err; echo "c4 dpall rc=$?"; cut -c1-250 gpurun_out/r2_bench_c4_n8_dpall.json
timeout 120 $TR --master-port 29524 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r2_bench_c1_n8.json 2> gpurun_out/r2_bench_c1_n8.err; echo "c1 n8 rc=$?"; cut -c1-250 gpurun_out/r2_bench_c1_n8.json
