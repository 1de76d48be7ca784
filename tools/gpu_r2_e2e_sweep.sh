for cfg in "--chunk 32 --streams 8" "--chunk 64 --streams 8" "--chunk 64 --streams 4" "--chunk 16 --streams 8" "--chunk 32 --streams 12" "--chunk 48 --streams 8"; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-grey $cfg 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['sections_pass']['ms_per_step'],1))"
done
