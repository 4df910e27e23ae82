#!/bin/bash
# Round 2, visit Z (1 GPU): coalesced first-pass stores of the envelope FFT -- parity tests + per-kernel times.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_preprocess.py -q -m gpu -x -k "envelope" 2>&1 | tail -3
bash scripts/gpu_r2_y.sh
