#!/bin/bash
# usage: scripts/run_scan.sh variant[:env-spec;env-spec...] ...   (one process per variant, strict timeout)
mkdir -p gpurun_out
for V in "$@"; do
  name=${V%%:*}; specs="-"
  case "$V" in *:*) specs="${V#*:}";; esac
  echo "=== $name"
  IFS=';' read -ra SP <<< "$specs"
  B200_LDU_LIB=multiregionfoam_b200/lib/variants/$name.so timeout ${SCAN_TIMEOUT:-90} python scripts/sweep_env_scan.py "${SP[@]}" 2>&1 | grep -v "^\[b200\]"
  echo "exit $?"
done
