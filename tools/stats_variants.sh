#!/bin/bash
cd "$(dirname "$0")/.."
cp multideggs_b200/csrc/libmdeggs.so /tmp/libmdeggs_orig.so
for v in orig s3 s4 s5 s6 orig; do
  if [ "$v" = orig ]; then cp /tmp/libmdeggs_orig.so multideggs_b200/csrc/libmdeggs.so; else cp tools/variants/libmdeggs_$v.so multideggs_b200/csrc/libmdeggs.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu --no-cfg5 --no-rlm 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), 'stats_stage_us', round(d['roofline']['stage_ms']['ms_stats']*1e3,1), 'eager_total', round(d['roofline']['stage_ms']['ms_total'],4))"
done
cp /tmp/libmdeggs_orig.so multideggs_b200/csrc/libmdeggs.so
