#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -25
