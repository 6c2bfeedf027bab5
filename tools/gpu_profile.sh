#!/bin/bash
# One gpurun call: the ncu evidence of a round (run from the repo root on a B200 box).  Each ncu pass follows a plain run of the
# same command that exited 0.  Raw reports stay on the box; their CSV exports come back in gpurun_out/.
#   tools/gpu_profile.sh <round-prefix>
R=${1:-r2}
WHAT=${2:-all}          # all | launches | kernels
mkdir -p gpurun_out
set -o pipefail
if [ "$WHAT" != "kernels" ]; then
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-legs"
$BENCH > gpurun_out/${R}_plain_bench.json 2> gpurun_out/${R}_plain_bench.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 285 -c 320 --csv --log-file gpurun_out/${R}_launches.csv $BENCH > gpurun_out/${R}_ncu_bench.log 2>&1
echo "launch list rc=$?"
fi
if [ "$WHAT" = "launches" ]; then exit 0; fi
python tools/prof_kernels.py 1 > gpurun_out/${R}_plain_kernels.log 2>&1 || { echo "plain prof_kernels failed"; tail -5 gpurun_out/${R}_plain_kernels.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"window_attn|attn_block|linear_tc|mlp_tc|cattn|frontend|detect_decode|nms_|add_layernorm|row_stats|stats_finalize" -c 80 -o /tmp/${R}_kernels -f python tools/prof_kernels.py 1 > gpurun_out/${R}_ncu_kernels.log 2>&1
echo "set full rc=$?"
ncu -i /tmp/${R}_kernels.ncu-rep --page raw --csv > gpurun_out/${R}_kernels_raw.csv
ncu -i /tmp/${R}_kernels.ncu-rep --page source --csv --kernel-name regex:window_attn_win8 > gpurun_out/${R}_win8_source.csv
ncu -i /tmp/${R}_kernels.ncu-rep --page source --csv --kernel-name regex:window_attn_flash > gpurun_out/${R}_flash_source.csv
ncu -i /tmp/${R}_kernels.ncu-rep --page source --csv --kernel-name regex:mlp_tc > gpurun_out/${R}_mlp_source.csv
ls -la gpurun_out/${R}_* | head -20
