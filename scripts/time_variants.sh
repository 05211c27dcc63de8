#!/bin/bash
# times sampler fwd / bwd of every libb200corr*.so variant next to the package (scripts/time_step.py)
for so in understanding_flow_robustness_b200/libb200corr*.so; do
  echo -n "$(basename $so) "
  B200CORR_LIB=$PWD/$so python scripts/time_step.py 2>&1 | tail -1
done
