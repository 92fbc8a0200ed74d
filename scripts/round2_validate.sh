#!/bin/bash
# First GPU call of round 2 (kept as a record; output: profiles/r2_round2_validate.log): validated everything written after the
# round-1 GPU budget was spent.  The FEAST_BAND_PIVOT / FEAST_RUN_EXPERIMENTAL / FEAST_RUN_EXPENSIVE switches it sets no longer exist:
# the pivoted band LU is the only banded path and the tests are part of the default GPU suite.
#   gpurun --timeout 1500 -- 'bash scripts/round2_validate.sh > gpurun_out/round2_validate.log 2>&1; tail -40 gpurun_out/round2_validate.log'
set -x
nvidia-smi -L
# 1. band LU with pivoting across block rows: parity tests of the banded path, then the backward error where the unpivoted one fails
FEAST_BAND_PIVOT=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "banded or C4 or butterfly" 2>&1 | tail -15
for mb in 200 400 500; do
  FEAST_BAND_PIVOT=1 FEAST_RUN_EXPENSIVE=1 FEAST_BAND_MB=$mb timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k backward_error_many 2>&1 | grep -E "banded solver|passed|failed|rror"
done
# 2. gated GPU tests: ifeast / nlfeast_it mirrors, mixed-precision COCG
FEAST_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ifeast or nlfeast_it" 2>&1 | tail -25
FEAST_RUN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mixed_prec" 2>&1 | tail -25
# 3. mixed-precision SpMM / COCG at the C2 shape (steady-state leg only)
timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --mixed-prec 2>&1 | tail -2
timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | tail -2
# 4. C4 at full size with the pivoted band LU (store=false: 8.2 GB of factors per node do not fit 24 nodes on one GPU)
FEAST_BAND_PIVOT=1 timeout 900 python scripts/c4_run.py --mb 500 --m0 64 --nodes 24 --r 0.006 --iter 4 --no-store 2>&1 | tail -15
