#!/bin/bash
# end-to-end rate of vb200_encode_dsp against chunk size and the ramped schedule (experiments)
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
: > gpurun_out/e2e_sweep.txt
for cb in 2048 4096 8192; do
  for r in 1 0; do
    VB200_CHUNK_BLOCKS=$cb VB200_CHUNK_RAMP=$r python bench.py --steps 8 --warmup 3 --no-extra --streams 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk',$cb,'ramp',$r,'e2e %.0f' % d['e2e']['value'],'resident %.0f' % d['value'], {k: round(v,2) for k,v in d['roofline']['kernel_ms'].items()})" | tee -a gpurun_out/e2e_sweep.txt
  done
done
