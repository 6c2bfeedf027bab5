#!/bin/bash
# Runs the GPU test groups in separate processes so that one faulting kernel cannot poison the rest.
mkdir -p gpurun_out
for grp in frontend linear_tc conv2d_nhwc patch_merge bias_act upsample add_layernorm window_attention swin_block cattn detect nms model; do
  timeout 600 python -m pytest tests -m gpu -q --tb=short -k "$grp" > gpurun_out/t_$grp.log 2>&1
  echo "== $grp: $(tail -1 gpurun_out/t_$grp.log)"
  grep -E "^(FAILED|ERROR)" gpurun_out/t_$grp.log | head -12
done
