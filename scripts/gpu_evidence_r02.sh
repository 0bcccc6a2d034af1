#!/bin/bash
# Round-2 profiler evidence at BASELINE configs[2] (batch 32 -> UNet batch 64), collected in ONE gpurun call:
#   1. ncu launch list (gpu__time_duration.sum, unmodified clocks) of steady-state step-graph replays of `bench.py --batch 32`
#   2. every kernel of one eager batch-32 UNet CFG step with duration + DRAM bytes + tensor-pipe / DRAM utilisation
#   3. `ncu --set full --import-source on` of the top kernels at batch-32 shapes (summary CSV exported on the box)
# Usage: bash scripts/gpu_evidence_r02.sh [outdir]
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv --log-file "$OUT/r02_launches_bench_b32.csv" \
    python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline > "$OUT/r02_bench_under_ncu_b32.log" 2>&1
echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --metrics "$M" --clock-control none --csv --log-file "$OUT/r02_step_launches_b32.csv" \
    python scripts/ncu_step.py 32 > "$OUT/r02_ncu_step_b32.log" 2>&1
echo "step launches rc=$?"
UDT_NCU_NB=64 UDT_NCU_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -o "$OUT/r02_full_b32" -f \
    python scripts/ncu_targets.py linear conv fmha ln gn conv8 > "$OUT/r02_ncu_full_b32.log" 2>&1
echo "full set rc=$?"
python scripts/ncu_summary.py "$OUT/r02_full_b32.ncu-rep" "$OUT/r02_ncu_full_summary_b32.csv"
ncu -i "$OUT/r02_full_b32.ncu-rep" --page details --csv > "$OUT/r02_ncu_full_details_b32.csv" 2>/dev/null
rm -f "$OUT/r02_full_b32.ncu-rep"      # 70+ MB: only its CSV exports travel back (gpurun_out is capped at 64 MiB)
ls -la "$OUT" | tail -12
