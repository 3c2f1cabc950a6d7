#!/bin/bash
# A/B timing of the specialiser's sign / phase rules on one B200 (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash scripts/ab_sign_rules.sh'
# Writes gpurun_out/ab_*.json: the same bench command with the rules on (default) and off
# (QBOT_B200_BRANCHY_SIGNS=1 QBOT_B200_PHASE_GREEDY=1 = the kernel set of profiles/r01_bench_1gpu_v11_11sweeps.json), twice each,
# interleaved, so that box-to-box and thermal differences cancel.
set -u
mkdir -p gpurun_out
for rep in 1 2; do
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_rules_on_$rep.json 2> gpurun_out/ab_rules_on_$rep.err
  QBOT_B200_BRANCHY_SIGNS=1 QBOT_B200_PHASE_GREEDY=1 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_rules_off_$rep.json 2> gpurun_out/ab_rules_off_$rep.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_rules_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'], 1), 'gates/s', round(d['ms_per_step'], 2), 'ms/step', round(d['roofline']['avg_launch_ms'], 3), 'ms/sweep', d['clocks'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
