#!/bin/bash
# Round 2: run_full_system at cfg3 with the reference's output cadence (101 stamps, 110 MB each), tmpfs and disk
mkdir -p gpurun_out
df -h /dev/shm /tmp | tail -2
timeout 600 python scripts/run_full_system_cfg3.py --dir /dev/shm --stamps 100 2>&1 | tail -1 | tee gpurun_out/run_full_system_cfg3_shm.json
timeout 600 python scripts/run_full_system_cfg3.py --dir /tmp --stamps 20 2>&1 | tail -1 | tee gpurun_out/run_full_system_cfg3_disk20.json
timeout 600 python scripts/run_full_system_cfg3.py --dir /dev/shm --stamps 1 2>&1 | tail -1 | tee gpurun_out/run_full_system_cfg3_1stamp.json
