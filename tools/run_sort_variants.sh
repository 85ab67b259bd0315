cd /root/repo
timeout 300 python tools/sort_probe.py 30 2>&1 | sed "s/^/[atomic-first] /"
B200_SORT_STABLE_FIRST=1 timeout 300 python tools/sort_probe.py 30 2>&1 | sed "s/^/[stable-first] /"
timeout 300 python tools/sort_probe.py 26 2>&1 | sed "s/^/[atomic-first] /"
timeout 900 python -m pytest tests/test_sort_gpu.py -x -q -m gpu 2>&1 | tail -3
