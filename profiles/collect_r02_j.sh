#!/bin/bash
tag=${1:-r02_j}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_harness.py -m gpu -q > $out/${tag}_pytest_trunk.log 2>&1; echo "pytest(trunk) rc=$?"; tail -4 $out/${tag}_pytest_trunk.log
timeout 200 python tests/gpu_probe_nms.py 2>&1 | tee $out/${tag}_nms_stages.log | tail -14
timeout 300 python tools/bench_detector.py --config faster --steps 10 --warmup 3 2>/dev/null | cut -c1-330
