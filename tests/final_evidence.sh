# Round-end evidence pass (run on the GPU box): GPU tests, both bench arms, ncu launch lists and full-set captures.
# .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): only their CSV pages come back.
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_final.log
python bench.py > gpurun_out/r01_bench_final.json 2> gpurun_out/r01_bench_final.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err; echo ref rc=$?
QASR_GRAPHS=0 python tests/ncu_step.py 2 > gpurun_out/ncu_step_plain.log 2>&1 && QASR_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 181 -c 181 --csv --log-file gpurun_out/r01_launches_final2.csv python tests/ncu_step.py 2 > /dev/null 2>&1; echo launches rc=$?
QASR_GRAPHS=0 ncu --set full --clock-control none --import-source on -s 181 -c 17 -o /tmp/r01_step_full2 -f python tests/ncu_step.py 2 > gpurun_out/ncu_step_full2.log 2>&1; echo full rc=$?
ncu -i /tmp/r01_step_full2.ncu-rep --page raw --csv > gpurun_out/r01_step_full2_raw.csv 2>/dev/null
python tests/ncu_prefill.py > gpurun_out/ncu_prefill_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_sm100|rmsnorm|qknorm|causal_attention|cast_rows" -s 19 -c 19 -o /tmp/r01_prefill_full2 -f python tests/ncu_prefill.py > gpurun_out/ncu_prefill2.log 2>&1; echo prefill rc=$?
ncu -i /tmp/r01_prefill_full2.ncu-rep --page raw --csv > gpurun_out/r01_prefill_full2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
