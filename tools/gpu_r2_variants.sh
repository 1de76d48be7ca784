#!/bin/bash
# A/B of build-time variants: for each flag set rebuild the library on the box and take a short headline line
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "${VARIANTS:-;-DI2S_CANNY1_MINB=6}"
for v in "${VARS[@]}"; do
  I2S_NVCC_FLAGS="$v" python -c "from img2sgf_b200 import build; build.build(force=True)" 2>&1 | tail -1
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extras --no-grey 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['sections']; print('variant [$v] value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'single', round(d['sections_pass']['ms_per_step'],2), {k: round(s[k]['ms_per_step'],2) for k in ('sobel_nms','sobel_nms_rgb','median','vote','edge_list','radius')})"
done | tee gpurun_out/r2_variants.txt
