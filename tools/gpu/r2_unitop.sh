#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests/test_gpu_ops.py -k "shuffle_unit" tests/test_gpu_modules.py > gpurun_out/t_unitop.log 2>&1
echo "rc=$?"; tail -n 12 gpurun_out/t_unitop.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
