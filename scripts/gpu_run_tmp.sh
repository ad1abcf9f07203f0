TAG=r1g BENCH_ARGS="--no-torch-eager" bash scripts/gpu_profile.sh 2>&1 | grep -v "^-\|^total" | tail -12
echo "=== full capture: layernorm ==="
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 8 -c 2 -f -o gpurun_out/r1g_prof_layernorm python scripts/gpu_variants_diag.py > gpurun_out/r1g_prof_layernorm.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
