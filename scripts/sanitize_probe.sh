#!/bin/bash
# synccheck probe: which CTAs / source lines report, with and without CTA pairs (tuning build, UDT_IGEMM_PAIR=0)
OUT=${1:-gpurun_out}
SEL='conv3x3_fused_skip or linear_matches_torch or pair_mode_ragged_rows'
summ() {  # distinct (kernel line, block) pairs
  grep -E "Barrier error|by thread|Device Frame: void|at udt" "$1" | sed 's/+0x[0-9a-f]*//' | paste - - - - 2>/dev/null | \
    sed -E 's/by thread \(([0-9]+),0,0\)/thr/; s/=========//g' | awk '{$1=$1};1' | sort | uniq -c | sort -rn | head -30
}
timeout 300 compute-sanitizer --tool synccheck --print-limit 100000 --error-exitcode 0 python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" > /tmp/sync_pair.log 2>&1
echo "== production build (pairs on):"; grep -E "ERROR SUMMARY|passed|failed" /tmp/sync_pair.log | tail -3
grep -o "in block ([0-9]*,0,0)" /tmp/sync_pair.log | sort | uniq -c | sort -rn | head -20
grep -o "udt_[a-z]*\.cu:[0-9]*" /tmp/sync_pair.log | sort | uniq -c | sort -rn | head
grep -o "Barrier is located at shared address 0x[0-9a-f]*" /tmp/sync_pair.log | sort | uniq -c | sort -rn | head
grep -o "Host Frame: test_[a-z_0-9]* in" /tmp/sync_pair.log | sort | uniq -c
UDT_TRACE=1 python -m udifftext_b200.build --force > /dev/null 2>&1
UDT_IGEMM_PAIR=0 timeout 300 compute-sanitizer --tool synccheck --print-limit 1000 --error-exitcode 0 python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" > /tmp/sync_nopair.log 2>&1
echo "== tuning build, UDT_IGEMM_PAIR=0:"; grep -E "ERROR SUMMARY|passed|failed" /tmp/sync_nopair.log | tail -3
grep -o "udt_[a-z]*\.cu:[0-9]*" /tmp/sync_nopair.log | sort | uniq -c | sort -rn | head
python -m udifftext_b200.build --force > /dev/null 2>&1
