B200SEP_KA=single timeout 200 python tools/power_probe.py 2 3 4 5 6 18 20 50 51 52 2>&1 | tail -12
B200SEP_KA=pair timeout 200 python tools/power_probe.py 2 3 4 52 2>&1 | tail -5
