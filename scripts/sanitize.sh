#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / cluster protocols: the kernel parity tests of the
# igemm (CTA pairs, split-K, ragged rows), the TS-form FMHA, every GroupNorm / LayerNorm schedule and the small glue kernels,
# one run per tool.  Usage (on the GPU box):  bash scripts/sanitize.sh [outdir]   -> <outdir>/sanitizer_<tool>.log
# Each tool is bounded by its own timeout; a tool that does not finish is reported as such, never silently dropped.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='linear_matches_torch or pair_mode_ragged_rows or splitk_linear or splitk_conv or conv3x3_fused_skip or fmha_matches_torch or groupnorm_schedules or layernorm_schedules or softmax_rows or rowsum_norm_split or mha_small_f32 or sampler_glue'
for TOOL in ${TOOLS:-memcheck synccheck racecheck initcheck}; do
  LOG="$OUT/sanitizer_${TOOL}.log"
  echo "== compute-sanitizer --tool $TOOL" > "$LOG"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" >> "$LOG" 2>&1
  echo "== exit code $? (124 = tool timed out)" >> "$LOG"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$LOG" | tail -4
done
