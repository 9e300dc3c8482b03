#!/bin/bash
# local helper (CPU container): rebuild, and only if the build is clean run the quick GPU check under gpurun
set -e
cd /root/repo
make -C /root/repo/rpmplv2_b200/csrc -j4 > /tmp/build.log 2>&1 || { grep -E "error" /tmp/build.log | head; echo BUILD FAILED; exit 1; }
grep -E " error" /tmp/build.log && { echo BUILD FAILED; exit 1; }
/usr/local/graft/bin/gpurun --timeout 1800 -- "bash profiles/quick_check.sh ${1:-quick}" 2>&1 | tail -${2:-7}
