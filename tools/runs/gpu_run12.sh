#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
