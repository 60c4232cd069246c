#!/bin/bash
for dbg in "$@"; do MSHGNN_STACK_DEBUG=$dbg timeout 60 python tools/stack_debug_probe.py 2>&1 | tail -3; echo "debug=$dbg rc=$?"; done
