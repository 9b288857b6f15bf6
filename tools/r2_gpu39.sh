#!/bin/bash
# round-2 GPU call 39 (2 GPUs): the whole GPU suite on the final tree, two-rank tests included
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/r2_g39_pytest.log 2>&1; tail -4 $O/r2_g39_pytest.log
