#!/bin/bash
# compute-sanitizer passes (run on a GPU box; minutes): the class pipeline, the cooperative kernel (TMA + mbarrier), the
# two-phase write path with its slot-ownership protocol.
set -x
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefix_tree and classes8"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alphabet and classes8 and 1"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_concurrency.py -x -q -m gpu -k "coop_kernel_short or two_phase or duplicate_hashes"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_concurrency.py -x -q -m gpu -k "coop_kernel_short"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged and coop"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_random_event_stream and rounds"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefix_tree and rounds2 and not lane and not spec and not nospec"
