#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r44_pytest.log 2>&1; tail -3 $O/r44_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r44_smoke.log 2>&1; tail -2 $O/r44_smoke.log
