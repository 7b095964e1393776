for v in default hints2 hints0; do
  for a in 2 4 8 16; do
    lib=multiregionfoam_b200/lib/libb200ldu.so; [ $v != default ] && lib=multiregionfoam_b200/lib/variants/$v.so
    B200_SWEEP_L2AHEAD=$a B200_LDU_LIB=$lib NO_STATS=1 python scripts/sweep_variants.py "$v-ahead$a" 2>&1 | grep C2
  done
done
