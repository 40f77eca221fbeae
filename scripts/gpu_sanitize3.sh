#!/bin/bash
# compute-sanitizer over the paths added in round 2: feature-regression / answer-head kernels, differentiable head
# forwards, the specialised GEMM epilogues (probe on ragged shapes), dropout, sampler transitions, packed step inputs
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_feat_qa.py -m gpu -x -q -k "visual_losses or answer_head or differentiable" > gpurun_out/sanitize3_heads.log 2>&1
echo "memcheck heads exit=$?"; tail -3 gpurun_out/sanitize3_heads.log
for s in "1000 776 328 3 0 0 651" "130 40 96 3 0 0 521" "300 64 768 3 0 0 520" "1000 776 328 3 0 0 1025" "777 3072 768 3 0 0 584" "1000 776 328 3 0 0 1024"; do
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 10 ./xlxmert_b200/lib/gemm_test $s > gpurun_out/sanitize3_probe.log 2>&1
  echo "memcheck probe [$s] exit=$?"; grep -E "ERROR SUMMARY| OK|FAIL" gpurun_out/sanitize3_probe.log | tail -2
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_dropout.py tests/test_sampler.py tests/test_inputs.py -m gpu -x -q -k "same_masks or transition or qa_and_feature or unpacked" > gpurun_out/sanitize3_misc.log 2>&1
echo "memcheck misc exit=$?"; tail -3 gpurun_out/sanitize3_misc.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_sampler.py tests/test_feat_qa.py -m gpu -x -q -k "transition or answer_head" > gpurun_out/sanitize3_race.log 2>&1
echo "racecheck exit=$?"; tail -3 gpurun_out/sanitize3_race.log
