cd $GRAFT_REPO_ROOT
for r in 0.5 0.46 0.42 0.38; do
echo "ratio $r"; B200SEP_SPLIT_RATIO=$r timeout 100 python tests/e2e_profile.py 2>&1 | tail -1
B200SEP_SPLIT_RATIO=$r timeout 100 python tests/split_time.py 2>&1 | tail -1 | cut -c1-80
done
