#!/bin/bash
for L in 0 278 371 556 1112; do
  if [ $L -eq 0 ]; then unset GPAR_KF_L; else export GPAR_KF_L=$L; fi
  echo "L=$L"; python tools/prof_kalman_grad.py 2>&1 | tail -1
done
