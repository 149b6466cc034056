#!/bin/bash
python tools/phase_times.py 2>&1 | grep -E "graph 0|head|stamps rel"
