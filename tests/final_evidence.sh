# Round-end evidence pass (run on the GPU box): GPU tests, both bench arms, ncu launch lists and full-set captures.
#   bash tests/final_evidence.sh r02
# .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): only their CSV pages come back.
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu_final.log
T0=$(date +%s); python bench.py > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err; echo bench rc=$? wall=$(( $(date +%s) - T0 ))s
T0=$(date +%s); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo ref rc=$? wall=$(( $(date +%s) - T0 ))s
QASR_GRAPHS=0 python tests/ncu_step.py 2 > gpurun_out/ncu_step_plain.log 2>&1; L=$(grep -o "launches per step: [0-9]*" gpurun_out/ncu_step_plain.log | grep -o "[0-9]*$"); echo "launches per step: $L"
[ -n "$L" ] && QASR_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -s $L -c $L --csv --log-file gpurun_out/${TAG}_launches_final.csv python tests/ncu_step.py 2 > /dev/null 2>&1; echo launches rc=$?
QASR_GRAPHS=0 ncu --set full --clock-control none --import-source on -s $L -c 17 -o /tmp/${TAG}_step_full2 -f python tests/ncu_step.py 2 > gpurun_out/ncu_step_full2.log 2>&1; echo full rc=$?
ncu -i /tmp/${TAG}_step_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full2_raw.csv 2>/dev/null
python tests/ncu_prefill.py > gpurun_out/ncu_prefill_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_sm100|rmsnorm|qknorm|causal_attention|cast_rows" -s 19 -c 19 -o /tmp/${TAG}_prefill_full2 -f python tests/ncu_prefill.py > gpurun_out/ncu_prefill2.log 2>&1; echo prefill rc=$?
ncu -i /tmp/${TAG}_prefill_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_prefill_full2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
