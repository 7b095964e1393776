#!/bin/bash
# one-box comparison of environment settings on bench workloads: scripts/env_scan.sh W:VAR=val[,VAR=val] ...   (W: alone = defaults)
for spec in "$@"; do
W=${spec%%:*}; E=${spec#*:}
tag=$(echo "$spec" | tr -c 'A-Za-z0-9\n' '_')
( IFS=','; for kv in $E; do [ -n "$kv" ] && export "$kv"; done
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/e_$tag.json 2>/tmp/e_$tag.err )
python - "$spec" "$tag" <<'PY'
import json,sys
spec,tag=sys.argv[1:3]
try:
    d=json.loads(open(f"/tmp/e_{tag}.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print(spec,f"{d['value']/1e9:.3f} G frac {d.get('solve_roofline_frac',0):.3f}"," ".join(f"{n}={k[n]['ms_per_launch']*1e3:.0f}us" for n in ("amul","sweep_fwd","sweep_bwd","vector") if n in k),flush=True)
except Exception as e: print(spec,"FAILED",e,open(f"/tmp/e_{tag}.err").read()[-600:])
PY
done
