cd $GRAFT_REPO_ROOT
timeout 300 python tools/dev_timeline.py 60 3 2>&1 | grep -v "^scan \|^stream scan\|^host" | tail -16
