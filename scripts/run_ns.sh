export B200_SWEEP_DEBUG=2
for cfg in "110 4" "220 6" "220 8"; do
  set -- $cfg
  echo "== smemKB $1 stages $2"
  B200_SWEEP_SMEM_KB=$1 B200_SWEEP_STAGES=$2 python scripts/sweep_hops.py 996 123 22 3 2>&1 | grep -A2 "dir +1" | cut -c1-150
  B200_SWEEP_SMEM_KB=$1 B200_SWEEP_STAGES=$2 python scripts/sweep_hops.py 2>&1 | grep -A2 "dir +1" | cut -c1-150
done
