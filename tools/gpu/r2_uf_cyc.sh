#!/bin/bash
export PYTHONUNBUFFERED=1
EXTRA=$((1<<24)) python tools/uf_phase_cycles.py 2>&1 | tail -3
EXTRA=$((1<<25)) python tools/uf_phase_cycles.py 2>&1 | tail -3
