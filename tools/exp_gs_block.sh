#!/bin/bash
# Round-2 experiment (one gpurun call, ~4 min): the layout of the compact gradient scratch.
#   BRS_GS_BLOCK = 8   sector-blocked (default): a row's 16 sectors on 16 lines / many L2 slices, 16 requests per REDG.128
#   BRS_GS_BLOCK = 32  line-blocked: 4 requests per REDG.128, a row over 4 lines 'capacity x 128 B' apart
#   BRS_GS_BLOCK = 128 (= D at the benchmark shape) row-major: 4 requests, a row in 4 adjacent lines (2 slices)
# For each: rebuild, MF parity tests, bench line (kernel times from the instrumented replay).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/exp_gs_block.sh > gpurun_out/exp_gs_block.log 2>&1'
set -u
cd "$(dirname "$0")/.."
for blk in 32 128 8; do
  echo "=== BRS_GS_BLOCK=$blk"
  BRS_NVCC_DEFINES="-DBRS_GS_BLOCK=$blk" python -m beta_recsys_b200.build --force > /dev/null || { echo "build failed"; continue; }
  timeout 200 python -m pytest tests/test_mf_gpu.py -m gpu -q -x 2>&1 | tail -1
  timeout 120 python bench.py --steps 1000 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.1f M/s  step %.1f us  prepass %.1f  fused %.1f  apply %.1f  frac %.3f' % (d['value']/1e6, d['ms_per_step']*1e3, r['prepass_kernel_ms']*1e3, r['kernel_ms']*1e3, r['apply_kernel_ms']*1e3, r['frac']))"
done
# the loop ends on the default layout: the in-tree library is the shipped one again
