#!/bin/bash
for d in 4 12; do echo "== PTTS_DBG=$d"; PTTS_DBG=$d timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | cut -c1-150; done
