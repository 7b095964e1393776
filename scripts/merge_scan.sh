#!/bin/bash
# sweep schedule with / without filling the last warp of a plane from a later plane: scripts/merge_scan.sh W:gap ...   (gap "off" = no merging)
for spec in "$@"; do
W=${spec%%:*}; G=${spec#*:}
if [ "$G" = off ]; then export B200_MERGE_SEQ=0; unset B200_MERGE_GAP; else export B200_MERGE_SEQ=1 B200_MERGE_GAP=$G; fi
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/m_$W$G.json 2>/tmp/m_$W$G.err
python - $W $G <<'PY'
import json,sys
W,M=sys.argv[1:3]
try:
    d=json.loads(open(f"/tmp/m_{W}{M}.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print(W,"gap",M,f"{d['value']/1e9:.3f} G frac {d['solve_roofline_frac']:.3f}"," ".join(f"{n}={k[n]['ms_per_launch']*1e3:.0f}us" for n in ("amul","sweep_fwd","sweep_bwd","vector")),flush=True)
except Exception as e: print(W,M,"FAILED",e,open(f"/tmp/m_{W}{M}.err").read()[-600:])
PY
done
