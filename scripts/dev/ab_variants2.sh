#!/bin/bash
# runs ab_tail.py for the default build and every grace-devel_b200/build/var_*.so (budget via $2)
LRS=${1:-20,17}
python scripts/dev/ab_tail.py $LRS default | cut -c1-190
for v in grace-devel_b200/build/var_*.so; do
  GRACE_B200_LIB=$PWD/$v timeout 200 python scripts/dev/ab_tail.py $LRS $(basename $v .so) | cut -c1-190
done
