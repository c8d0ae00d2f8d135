#!/bin/bash
RPGO_TRACE=1 timeout 300 python tools/e2e_probe.py 50000 2>&1 | tail -40
