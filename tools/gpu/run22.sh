cd $GRAFT_REPO_ROOT
timeout 95 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q -m gpu --maxfail=4 2>&1 | tail -5
