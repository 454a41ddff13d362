#!/bin/bash
for d in 256 1024 2048 4096; do echo "== PTTS_DBG=$d (tile pieces $((d/256)))"; PTTS_DBG=$d timeout -s KILL 100 python -u tools/profile_step.py 100 2>&1 | tail -13 | cut -c1-150 | grep -v "^phase\|attn\|embed"; done
