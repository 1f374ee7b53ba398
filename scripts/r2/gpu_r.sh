#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 1200 python -m pytest tests/test_dataset.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
