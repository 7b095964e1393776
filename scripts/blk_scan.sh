#!/bin/bash
# compare differently compiled builds of the block sweeps on one box: scripts/blk_scan.sh <workload> name[@gateSlack]...
W=$1; shift
for spec in "$@"; do
  v=${spec%%@*}; slack=0
  case "$spec" in *@*) slack=${spec#*@};; esac
  B200_BLK_GATE_SLACK=$slack B200_LDU_LIB=$PWD/multiregionfoam_b200/lib/variants/$v.so timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/blk_$v.json 2>/tmp/blk_$v.err
  python - "$v" "$slack" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"/tmp/blk_{v}.json").read().strip().splitlines()[-1])
    k = d["kernels"]
    print(v, "slack", sys.argv[2], f"{d['value']/1e9:.3f} G", " ".join(f"{n}={k[n]['ms_per_launch']:.3f}ms" for n in ("amul", "sweep_fwd", "sweep_bwd") if n in k), flush=True)
except Exception as e:
    print(v, "FAILED", e, open(f"/tmp/blk_{v}.err").read()[-400:])
PY
done
