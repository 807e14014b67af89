#!/bin/bash
# quick GPU check of a change: the whole -m gpu suite, one bench line (no extras), psy phase timing
python -m pytest tests -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
python bench.py --no-extra --streams 0 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/q_bench.json') if l.startswith('{')][0]
print('value', d['value'], 'e2e', d['e2e']['value'], 'phaseA', d['roofline'].get('phaseA_only_blocks_per_s'))
print(d['roofline']['kernel_ms'])
PY
python tools/phase_timing.py > gpurun_out/q_phase_timing.txt 2>&1; tail -16 gpurun_out/q_phase_timing.txt
