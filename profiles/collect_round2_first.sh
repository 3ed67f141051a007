#!/bin/bash
# (historical: the CTA-pair forward and the BAGS_TEST_EXPERIMENTAL gates it exercises were removed later in round 2)
# First GPU call of the next round (1 GPU, ~3 min): state of the tree + the experiments prepared at the end of round 1.
#   gpurun --timeout 1200 -- 'bash profiles/collect_round2_first.sh r02_v0'
tag=${1:-r02_v0}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench.json
# 1. CTA-pair forward (BAGS_FWD_PAIR=1): parity vs the default fused forward + kernel timing A/B
BAGS_DBG_OCC=1 timeout 300 python tests/gpu_probe.py --case fwdpair.bf16 > $out/${tag}_fwdpair.log 2>&1; echo "fwdpair rc=$?"; grep -m1 'max active clusters' $out/${tag}_fwdpair.log; tail -3 $out/${tag}_fwdpair.log | cut -c1-1500
# 2. whole step with the pair forward
for v in 0 1 0 1; do echo -n "BAGS_FWD_PAIR=$v "; BAGS_FWD_PAIR=$v timeout 200 python bench.py --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_fwdpair_step_ab.log
# 3. library yardstick: the three plain GEMMs through cuBLAS from a CUDA graph (gpu_probe 'timing': cublas_three_gemms_graph_us)
timeout 200 python tests/gpu_probe.py --case timing > $out/${tag}_timing.log 2>&1; grep RESULT $out/${tag}_timing.log | cut -c1-900
# 4. ticket-based preparation jobs in the merged backward (BAGS_BWD_TICKET=1): parity suite under the flag + step A/B
BAGS_BWD_TICKET=1 timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_ticket.log 2>&1; echo "pytest (ticket) rc=$?"; tail -2 $out/${tag}_pytest_ticket.log
for v in 0 1 0 1; do echo -n "BAGS_BWD_TICKET=$v "; BAGS_BWD_TICKET=$v timeout 200 python bench.py --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_ticket_step_ab.log
# 5. detector harness: frozen torchvision trunk + BAGS head(s), synthetic 1333x800 images
timeout 300 python tools/bench_detector.py --stages 1 --steps 10 --warmup 3 > $out/${tag}_detector_s1.json 2> $out/${tag}_detector_s1.err; echo "detector s1 rc=$?"; cat $out/${tag}_detector_s1.json; tail -3 $out/${tag}_detector_s1.err
timeout 300 python tools/bench_detector.py --stages 3 --steps 10 --warmup 3 > $out/${tag}_detector_s3.json 2> $out/${tag}_detector_s3.err; echo "detector s3 rc=$?"; cat $out/${tag}_detector_s3.json; tail -3 $out/${tag}_detector_s3.err
# 6. experimental one-launch class-aware NMS against the oracle
BAGS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_nms.py -q -m gpu > $out/${tag}_pytest_nms.log 2>&1; echo "nms rc=$?"; tail -3 $out/${tag}_pytest_nms.log
# 7. experimental reweight head variant (fp32 per-RoI weights through bags_fwd_w / bags_group_ce_w)
BAGS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_reweight.py -q -m gpu > $out/${tag}_pytest_reweight.log 2>&1; echo "reweight rc=$?"; tail -3 $out/${tag}_pytest_reweight.log
