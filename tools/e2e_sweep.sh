for cb in 1024 2048 4096 8192 16384; do
  VB200_CHUNK_BLOCKS=$cb python bench.py --steps 3 --warmup 3 --no-extra --streams 0 2> gpurun_out/r2f_e2e_$cb.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk',$cb,'e2e',d['e2e']['value'],'resident',d['value'])" >> gpurun_out/r2f_e2e.txt
done
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
