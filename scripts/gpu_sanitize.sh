#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02y}"
timeout 120 python tests/dev_sanitize.py > $O/${T}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/${T}_sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/dev_sanitize.py > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $O/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/dev_sanitize.py attention score layernorm > $O/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 $O/${T}_racecheck.log
