#!/bin/bash
# A/B timing of how the bulk copies of the next tile are issued by the specialised sweeps (one B200, through gpurun):
#   default              one elected lane per warp, warp-uniform addresses (straight-line UBLKCP)
#   QBOT_B200_LANE_ISSUE=1   one copy per lane (ptxas serialises the lanes in a waterfall loop; the kernel set of
#                            profiles/r02_qj_head_ncu_summary.txt)
#   QBOT_B200_JIT_L2_LATE=1  default + L2 prefetch of the late half together with the early half
# interleaved so that box-to-box and thermal differences cancel; then the specialised-sweep GPU tests, then the ncu
# passes of the new kernel set (per-launch DRAM bytes + durations of all 11 sweeps, --set full + source view of the heaviest).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
timeout 120 $B > gpurun_out/ab_issue_uniform_1.json 2> gpurun_out/ab_issue_uniform_1.err
QBOT_B200_LANE_ISSUE=1 timeout 90 $B --no-parity > gpurun_out/ab_issue_lanes_1.json 2> gpurun_out/ab_issue_lanes_1.err
QBOT_B200_JIT_L2_LATE=1 timeout 90 $B --no-parity > gpurun_out/ab_issue_l2late_1.json 2> gpurun_out/ab_issue_l2late_1.err
timeout 90 $B --no-parity > gpurun_out/ab_issue_uniform_2.json 2> gpurun_out/ab_issue_uniform_2.err
QBOT_B200_LANE_ISSUE=1 timeout 90 $B --no-parity > gpurun_out/ab_issue_lanes_2.json 2> gpurun_out/ab_issue_lanes_2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_issue_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'], 1), 'gates/s', round(d['ms_per_step'], 2), 'ms/step', round(d['roofline']['avg_launch_ms'], 3), 'ms/sweep',
              d['roofline'].get('kernel_set'), d.get('parity_check', {}).get('status', d.get('parity_check')), d['clocks'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
timeout 150 python -m pytest tests/test_gpu_jit.py -x -q > gpurun_out/ab_issue_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_issue_pytest.log
tail -3 gpurun_out/ab_issue_pytest.log
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:qj_kernel -c 22 --csv \
    --log-file gpurun_out/r02_qj_uniform_launches.csv python scripts/traffic_capture.py run > gpurun_out/ab_issue_ncu1.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:qj_kernel -s 11 -c 1 -f -o gpurun_out/r02_qj_uniform \
    python scripts/traffic_capture.py run > gpurun_out/ab_issue_ncu2.log 2>&1
ncu -i gpurun_out/r02_qj_uniform.ncu-rep --page raw --csv > gpurun_out/r02_qj_uniform_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_qj_uniform.ncu-rep --page source --csv > gpurun_out/r02_qj_uniform_sweep0_source.csv 2>/dev/null
rm -f gpurun_out/r02_qj_uniform.ncu-rep
tail -2 gpurun_out/ab_issue_ncu1.log gpurun_out/ab_issue_ncu2.log
