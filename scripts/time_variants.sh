#!/bin/bash
# runs a timing script (default scripts/time_step.py) against every libb200corr*.so variant next to the package
S=${1:-scripts/time_step.py}
for so in understanding_flow_robustness_b200/libb200corr*.so; do
  echo -n "$(basename $so) "
  B200CORR_LIB=$PWD/$so python $S 2>&1 | tail -1
done
