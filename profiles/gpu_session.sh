#!/bin/bash
# One GPU session (gpurun -- 'bash profiles/gpu_session.sh [tag]'): smoke, full GPU suite, default bench, launch list and a
# full ncu capture of the step kernels.  Everything that must come back goes to gpurun_out/ and stays well under the 64 MiB
# merge limit: the .ncu-rep (65-70 MB with sources) is written to /tmp and only its CSV exports are kept.
# Variants for A/B runs:  PSL_LIB=/path/to/other.so (another build of the same sources, point_slam_b200/_lib.py),
# PSL_W16=1 (16-worker-warp forward experiment), PSL_EXPERIMENTAL=1 (opt-in tests), PSL_OVERLAP=0, PSL_TC=0, PSL_TC_BWD=0.
TAG=${1:-session}
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -rxX ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR|XPASS|XFAIL)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -12
( time timeout 600 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/{sys.argv[1]}_bench.json').read().splitlines() if l.startswith('{')][-1])
print(round(d['ms_per_step'], 2), 'ms/step; e2e', round(d['e2e']['ms_per_step'], 2), '; roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 4))
print(d['kernel_ms_per_step'], d.get('map_maintenance'))
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python profiles/prof_step.py 1 2 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"k_color_fwd_tc|k_color_bwd_tc|k_wgrad_tc|k_decode_bwd|k_decode_fwd|k_knn|k_scatter_segments|k_adam_rows|k_frustum|k_add_probe" \
    -f -o /tmp/${TAG}_full python profiles/prof_step.py 1 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv -k regex:k_color_bwd_tc -c 1 > gpurun_out/${TAG}_source_color_bwd.csv 2>/dev/null
du -sh gpurun_out
