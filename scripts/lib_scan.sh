#!/bin/bash
# one-box comparison of differently compiled libraries on a bench workload: scripts/lib_scan.sh <workload> name...
W=$1; shift
for v in "$@"; do
  B200_LDU_LIB=$PWD/multiregionfoam_b200/lib/variants/$v.so timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/l_$v.json 2>/tmp/l_$v.err
  python - "$v" "$W" <<'PY'
import json, sys
v, W = sys.argv[1:3]
try:
    d = json.loads(open(f"/tmp/l_{v}.json").read().strip().splitlines()[-1]); k = d["kernels"]
    print(W, v, f"{d['value']/1e9:.3f} G frac {d.get('solve_roofline_frac', 0):.3f}", " ".join(f"{n}={k[n]['ms_per_launch']*1e3:.0f}us" for n in k if n in ("amul", "sweep_fwd", "sweep_bwd", "vector")), flush=True)
except Exception as e:
    print(v, "FAILED", e, open(f"/tmp/l_{v}.err").read()[-400:])
PY
done
