cd $GRAFT_REPO_ROOT
timeout 300 python tools/dev_timeline.py 60 4 2>&1 | grep "^wave\|^chunk\|ms_total" | tail -14
