#!/bin/bash
# scratch: the test subset / A-B of the change being worked on
timeout 70 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
