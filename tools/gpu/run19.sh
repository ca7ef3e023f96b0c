cd $GRAFT_REPO_ROOT
timeout 150 python tools/gpu/diag19.py 2>&1 | tail -14
