cd $GRAFT_REPO_ROOT
timeout 120 python tests/gpu_diag.py e2e 2>&1 | grep "n_src=3 fp16 \|FAILED\|rror\|abort" | head -3
for ms in 1 2 4 8; do echo "merge_split $ms"; B200SEP_MERGE_SPLIT=$ms B200SEP_SPLIT=1 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200; done
for i in 1 2; do timeout 100 python tests/split_time.py 2>&1 | tail -1; B200SEP_MERGE_SPLIT=1 timeout 100 python tests/split_time.py 2>&1 | tail -1; done
