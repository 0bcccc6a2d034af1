#!/bin/bash
# One GPU-box pass over everything a change to a kernel must keep green: the -m gpu suite, smoke(), the bench line and the
# kernel-vs-library table.  Outputs land in gpurun_out/ (copy what should be kept into profiles/).
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_validate.sh [tag]'
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 900 gpurun_out/${tag}_bench_n1.json
python scripts/vs_library.py 8 > gpurun_out/${tag}_vs_library_b4.jsonl 2>/dev/null
python scripts/vs_library.py 64 > gpurun_out/${tag}_vs_library_b32.jsonl 2>/dev/null
grep fmha gpurun_out/${tag}_vs_library_b4.jsonl gpurun_out/${tag}_vs_library_b32.jsonl
