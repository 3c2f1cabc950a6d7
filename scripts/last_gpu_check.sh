#!/bin/bash
# Last GPU call of round 2 (3.8 GPU-minutes left): is HEAD's way of issuing the next tile's bulk copies
# (one elected lane per warp, uniform addresses; stage 0 waits per half) correct and faster than one copy per lane?
set -u
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
timeout 100 $B > gpurun_out/ab_issue_uniform_1.json 2> gpurun_out/ab_issue_uniform_1.err; echo "uniform_1 rc=$?"
tail -c 600 gpurun_out/ab_issue_uniform_1.err
QBOT_B200_LANE_ISSUE=1 timeout 60 $B --no-parity > gpurun_out/ab_issue_lanes_1.json 2> gpurun_out/ab_issue_lanes_1.err; echo "lanes_1 rc=$?"
timeout 120 python -m pytest tests/test_gpu_jit.py -x -q > gpurun_out/ab_issue_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/ab_issue_pytest.log
tail -3 gpurun_out/ab_issue_pytest.log
timeout 60 $B --no-parity > gpurun_out/ab_issue_uniform_2.json 2> gpurun_out/ab_issue_uniform_2.err; echo "uniform_2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_issue_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'], 1), 'gates/s', round(d['ms_per_step'], 2), 'ms/step', d['roofline'].get('kernel_set'),
              (d.get('parity_check') or {}).get('status'), d['clocks'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
