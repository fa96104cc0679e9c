#!/bin/bash
timeout 600 python scripts/adv_grad_diag.py 2>&1 | grep -v Warning | tail -12 | cut -c1-330
