cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6; do
timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-60
done
echo "--- PDL off"
for i in 1 2 3 4 5 6; do
B200SEP_PDL=0 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-60
done
nvidia-smi --query-gpu=name,ecc.errors.uncorrected.volatile.total --format=csv
