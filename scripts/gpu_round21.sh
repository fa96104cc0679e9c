#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/host_time_probe.py 160 2>&1 | tail -3
timeout 300 python scripts/host_time_probe.py 64 2>&1 | tail -3
