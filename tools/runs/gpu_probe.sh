#!/bin/bash
python tools/phase_times.py 2>&1 | grep -E "probe|stamps rel"
