#!/bin/bash
# compute-sanitizer on a small parity subset (ONE tool per gpurun call: tools/sanitizer.sh memcheck | racecheck | initcheck | synccheck).
# The subset covers both count engines (tcgen05 GEMM with TMA + mbarriers, inverted-index atomics), the item-space kernels, the select,
# the blends, a song partition and the join of partitions, on c1-sized and smaller shapes.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='fixture_4_3 or golden_small or config_c1 or edge_cases or (synthetic_parity and 300)'
python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_plain.log 2>&1 || { tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 99 --log-file gpurun_out/r02_sanitizer_$TOOL.log \
  python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/r02_sanitizer_${TOOL}_pytest.log 2>&1
echo "parity subset under $TOOL: rc=$?"
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 99 --log-file gpurun_out/r02_sanitizer_${TOOL}_window.log \
  python -m pytest tests/test_gpu_window.py -x -q -k "shape0 or argument_checks" > gpurun_out/r02_sanitizer_${TOOL}_window_pytest.log 2>&1
echo "window subset under $TOOL: rc=$?"
tail -3 gpurun_out/r02_sanitizer_${TOOL}_pytest.log gpurun_out/r02_sanitizer_${TOOL}_window_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$TOOL.log gpurun_out/r02_sanitizer_${TOOL}_window.log
