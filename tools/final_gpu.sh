#!/bin/bash
# final measurement pass of the round (run under gpurun on one B200)
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
# launch list (cold-cache, serialised: shares only) and the full capture of one un-split 24000-block step
VB200_SPLIT=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 12 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --blocks 24000 --no-extra --streams 0 > gpurun_out/r2_launches.log 2>&1
VB200_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"k_phaseA_transform|k_phaseA_psy3|k_floor1_fit|k_floor1_render|k_cqn_fast" -s 15 -c 5 -o gpurun_out/r2_chain python bench.py --steps 1 --warmup 3 --blocks 24000 --no-extra --streams 0 > gpurun_out/r2_chain_ncu.log 2>&1
python tools/phase_timing.py > gpurun_out/r2_phase_timing.txt 2>&1
python tools/pcie_bw.py > gpurun_out/r2_pcie_bw.json 2>&1

