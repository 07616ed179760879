#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/eigh_lab.py 2>&1 | tee gpurun_out/eigh_lab.log
timeout 300 python tools/sanitize.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain walk rc=$?"; tail -2 gpurun_out/sanitize_plain.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|SANITIZE_WALK_OK|ok$|Error" gpurun_out/sanitize_memcheck.log | head -20
