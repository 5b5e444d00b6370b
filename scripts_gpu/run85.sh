cd $GRAFT_REPO_ROOT
for i in 1 2 3 4; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
for i in 1 2 3 4 5 6 7 8; do timeout 100 python tests/split_time.py 2>&1 | tail -1 | cut -c1-90; done
for i in 1 2 3; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['clocks']['samples'])"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
