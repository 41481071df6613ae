#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multirank.py -q -k "29611 or 29621 or 29613" --timeout 200 2>&1 | tail -3
