#!/bin/bash
# compute-sanitizer over the parity tests of the kernels (one B200); summaries go to gpurun_out/r2_sanitizer.txt
out=gpurun_out/r2_sanitizer.txt
: > $out
K1="44k_stereo_q5 and (encode_dsp or encode_streams or plan_blocks or floor1 or phaseA or envelope or residue or decode or synthesis or mdct or drft)"
echo "== memcheck: pytest -m gpu -k \"$K1\"" >> $out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K1" > gpurun_out/san_mem.log 2>&1
echo "exit $?" >> $out; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem.log | tail -3 >> $out
K2="44k_stereo_q5 and (encode_dsp_one_call or encode_streams or floor1 or phaseA_vs or envelope or residue_classify or decode_dsp or mdct_forward)"
echo "== racecheck: pytest -m gpu -k \"$K2\"" >> $out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K2" > gpurun_out/san_race.log 2>&1
echo "exit $?" >> $out; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race.log | tail -3 >> $out
cat $out
