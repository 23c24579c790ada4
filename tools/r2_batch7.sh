#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_cabi_gpu.py -m gpu -q -x -s -k "fallback or device_buil" 2>&1 | tail -30
