#!/bin/bash
# A/B: kernel G of all parts on one stream (serial) x its grid x parts, with the deeper hash staging (lib) and the old one (lib_exp/s2)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 600 python scripts/ab_step.py 10000000 1048576 base KVIDX_GROUP_SERIAL=1,KVIDX_GROUP_SERIAL_GRID=1 KVIDX_GROUP_SERIAL=1,KVIDX_GROUP_SERIAL_GRID=2 KVIDX_GROUP_SERIAL=1,KVIDX_GROUP_SERIAL_GRID=3 \
   KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=12 KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=16 KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=6 KVIDX_ROUNDS_PARTS=12 > $O/r12_ab_s4.txt 2>&1
cat $O/r12_ab_s4.txt
KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/s2/libkvidx.so timeout -s KILL 300 python scripts/ab_step.py 10000000 1048576 base KVIDX_GROUP_SERIAL=1 > $O/r12_ab_s2.txt 2>&1
cat $O/r12_ab_s2.txt
KVIDX_GROUP_SERIAL=1 timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r12_timeline_serial.json > $O/r12_timeline_serial.out 2>&1; tail -40 $O/r12_timeline_serial.out
