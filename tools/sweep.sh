#!/bin/bash
# BASELINE.json configs 3, 4, 5 on one GPU: JSON lines into gpurun_out/sweep.jsonl
out=gpurun_out/sweep.jsonl; : > $out
for b in 256 1024 4096 8192 16384 65536; do
  timeout 600 python bench.py --steps 3 --warmup 3 --batch $b --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --workload cstr --batch 4096 >> $out 2>> gpurun_out/sweep.err
timeout 900 python bench.py --steps 2 --warmup 3 --workload kite --batch 1024 >> $out 2>> gpurun_out/sweep.err
timeout 600 python bench.py --steps 3 --warmup 3 --sqp-max-iter 10 --ls-max-iter 10 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:60], '| it/s %.0f' % d['value'], '| e2e %.0f' % d['e2e']['value'], '| ms/step %.1f' % d['ms_per_step'], '| solved %.4f' % d['solved_fraction'], '| cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
