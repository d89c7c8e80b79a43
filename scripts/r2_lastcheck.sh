#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 150 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-yardstick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
