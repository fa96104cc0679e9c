#!/bin/bash
# scratch: the test subset / A-B of the change being worked on (edit freely)
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-200
