#!/bin/bash
# last validation of the round: full GPU suite, smoke, sanitizer passes, a long randomised soak
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r38_pytest.log 2>&1; tail -3 $O/r38_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r38_smoke.log 2>&1; tail -2 $O/r38_smoke.log
timeout -s KILL 900 bash scripts/sanitize.sh > $O/r38_sanitize.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|failed" $O/r38_sanitize.log
timeout -s KILL 900 python scripts/soak.py 1200 20000 > $O/r38_soak.log 2>&1; tail -2 $O/r38_soak.log
