cd $GRAFT_REPO_ROOT
for s in 1 2 3 4; do B200SEP_SPLIT=$s timeout 100 python tests/split_time.py 2>&1 | tail -1; done
timeout 100 python tests/split_time.py 2>&1 | tail -1
B200SEP_SPLIT=2 B200SEP_PDL=1 timeout 100 python tests/split_time.py 2>&1 | tail -1
B200SEP_SPLIT=1 timeout 100 python tests/split_time.py 256 2>&1 | tail -1
B200SEP_SPLIT=2 timeout 100 python tests/split_time.py 256 2>&1 | tail -1
B200SEP_SPLIT=4 timeout 100 python tests/split_time.py 256 2>&1 | tail -1
