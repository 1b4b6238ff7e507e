#!/bin/bash
# bench.py with the N3 leg (pipeline.n3_extraction), short form: no e2e / cpu / parity legs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 110 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > $O/c40_bench.json 2> $O/c40_bench.err
echo "bench rc=$?" > $O/c40_box.log
cat $O/c40_box.log; tail -3 $O/c40_bench.err
python - <<PY
import json
for line in open('$O/c40_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged')})
        print(d['pipeline'].get('n3_extraction'))
        print({k:v for k,v in d['pipeline'].get('n4_compartments',{}).items() if k in ('wall_ms','good_bins','error')})
PY
