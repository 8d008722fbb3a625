#!/usr/bin/env bash
# compute-sanitizer passes over the kernels with hand-rolled synchronisation (run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# ).  memcheck + synccheck on the tcgen05 / mbarrier / TMA kernels (small-shape operator tests), racecheck on the SIMT
# kernels that share memory through __shared__ (racecheck does not model the async proxy of TMA / tcgen05).
# Each pass is time-boxed (PFB_SANITIZE_TIMEOUT seconds, default 300): a cut-off pass reports what it covered.
set -x
T=${PFB_SANITIZE_TIMEOUT:-300}
export PFB_CUDA_GRAPH=0
UMMA="first_conv or flow_conv7x7 or conv_umma or corr_volume_umma or gru_epilogues"
SIMT="lookup_radius or instance_norm or forward_interpolate or upsample or conv2d_vs_torch"
run() { # tool, selection
  timeout "$T" compute-sanitizer --tool "$1" --error-exitcode 1 --launch-timeout 120 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_umma.py -m gpu -x -q -k "$2"
  rc=$?
  if [ $rc -eq 0 ]; then echo "SANITIZER $1 [$2]: CLEAN"; elif [ $rc -eq 124 ]; then echo "SANITIZER $1 [$2]: TIME-BOXED (no error before the cut-off)"; else echo "SANITIZER $1 [$2]: FAILED rc=$rc"; fi
}
# round-2 kernels: tiled volume / lookup, on-the-fly correlation on the tensor cores (tests/test_gpu_corr_block.py; the 1020-item case is too long under the tool)
run2() {
  timeout "$T" compute-sanitizer --tool "$1" --error-exitcode 1 --launch-timeout 120 \
    python -m pytest tests/test_gpu_corr_block.py -m gpu -x -q -k "(onthefly_tensor_core or tiled) and not 135"
  rc=$?
  if [ $rc -eq 0 ]; then echo "SANITIZER $1 [round-2 corr kernels]: CLEAN"; elif [ $rc -eq 124 ]; then echo "SANITIZER $1 [round-2 corr kernels]: TIME-BOXED (no error before the cut-off)"; else echo "SANITIZER $1 [round-2 corr kernels]: FAILED rc=$rc"; fi
}
if [ "${PFB_SANITIZE_ONLY_R2:-0}" = "1" ]; then run2 memcheck; run2 synccheck; exit 0; fi
if [ "${PFB_SANITIZE_ONLY_SIMT:-0}" = "1" ]; then run racecheck "$SIMT"; run memcheck "$SIMT"; exit 0; fi
run2 memcheck
run2 synccheck
run memcheck "$UMMA"
run synccheck "$UMMA"
run racecheck "$SIMT"
run memcheck "$SIMT"
