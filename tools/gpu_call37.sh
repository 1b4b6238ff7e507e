#!/bin/bash
# N3 extraction side: tb_region_cells through the goldens of the unmodified reference, against the
# oracle on a larger matrix, and inside the kernel coverage script (default and checked build)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 240 python -m pytest tests/test_extraction.py tests/test_gpu_coverage.py -m gpu -q --maxfail=6 > $O/c37_tests.log 2>&1
echo "tests rc=$?" > $O/c37_box.log
tail -25 $O/c37_tests.log; cat $O/c37_box.log
