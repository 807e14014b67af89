#!/bin/bash
# compute-sanitizer over the whole -m gpu suite (one B200); summaries go to gpurun_out/r2_sanitizer.txt
out=gpurun_out/r2_sanitizer.txt
: > $out
echo "== memcheck: compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q" >> $out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests -m gpu -q > gpurun_out/san_mem.log 2>&1
echo "exit $?" >> $out; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem.log | tail -3 >> $out
echo "== racecheck: compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q" >> $out
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests -m gpu -q > gpurun_out/san_race.log 2>&1
echo "exit $?" >> $out; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race.log | tail -3 >> $out
echo "== synccheck: compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 44k_stereo_q5" >> $out
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -k 44k_stereo_q5 > gpurun_out/san_sync.log 2>&1
echo "exit $?" >> $out; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_sync.log | tail -3 >> $out
cat $out
