#!/bin/bash
timeout -s KILL 120 python tools/profile_stage.py 2>&1 | tail -8
