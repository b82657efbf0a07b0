#!/bin/bash
# compute-sanitizer passes over the class-pipeline parity tests (run on a GPU box; minutes).
set -x
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefix_tree and classes8"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alphabet and classes8 and 1"
