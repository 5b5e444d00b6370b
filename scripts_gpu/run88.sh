cd $GRAFT_REPO_ROOT
timeout 100 python tests/sanitize_target.py 2>&1 | tail -1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python tests/sanitize_target.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|^ok" gpurun_out/sanitizer_memcheck.log | head -12
