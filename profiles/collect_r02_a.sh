#!/bin/bash
# Round 2, 1-GPU call A: parity suite with the N = 4096 / MT-boundary / graphed-step oracle tests and the un-gated NMS /
# reweight tests; bench line with the new legs; ticketed vs static preparation jobs; timelines of the current kernels.
tag=${1:-r02_a}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"; cut -c1-1200 $out/${tag}_bench_reference.json
for v in 1 0 1 0; do echo -n "BAGS_BWD_TICKET=$v "; BAGS_BWD_TICKET=$v timeout 200 python bench.py --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_ticket_step_ab.log
timeout 300 python tests/gpu_probe.py --case timeline > $out/${tag}_timeline.log 2>&1; echo "timeline rc=$?"; grep RESULT $out/${tag}_timeline.log | cut -c1-2500
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
