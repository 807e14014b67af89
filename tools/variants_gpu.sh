#!/bin/bash
# A/B of library variants built under vorbis_b200/_variants (experiments only)
cp vorbis_b200/libvorbis_b200.so /tmp/cur.so
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for v in cur prev cur prev; do
  if [ $v = cur ]; then cp /tmp/cur.so vorbis_b200/libvorbis_b200.so; else cp vorbis_b200/_variants/$v.so vorbis_b200/libvorbis_b200.so; fi
  python bench.py --no-extra --streams 0 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$v', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms %.3f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()})"
done
cp /tmp/cur.so vorbis_b200/libvorbis_b200.so
