#!/usr/bin/env bash
# compute-sanitizer passes over the kernels with hand-rolled synchronisation (run on a GPU box, e.g.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# ).  memcheck + racecheck on the small-shape operator tests; synccheck on the tcgen05 / mbarrier kernels.
# Not run in round 1 (the GPU budget went to measurement); first thing to run in round 2.
set -x
SEL="first_conv or flow_conv7x7 or conv_umma or corr_volume_umma or instance_norm or forward_interpolate"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 120 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_umma.py -m gpu -x -q -k "$SEL" || echo "SANITIZER $tool: FAILED"
done
