VB200_PSY_V4=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phaseA" > gpurun_out/r2p_pytest_v4.log 2>&1
VB200_PSY_V4=1 timeout 200 python tools/phase_timing.py > gpurun_out/r2p_phase_v4.txt 2>&1
timeout 200 python tools/phase_timing.py > gpurun_out/r2p_phase_v3.txt 2>&1
VB200MS_PROFILE=1 timeout 600 python tools/dropin_throughput.py > gpurun_out/r2p_dropin.json 2> gpurun_out/r2p_dropin.err
