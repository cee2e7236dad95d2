for v in default tile32x4; do
  if [ "$v" = default ]; then unset C2RT_LIB_DIR; else export C2RT_LIB_DIR=build_variants/$v; fi
  for rep in 1 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 40 --warmup 3 --no-scaling-target > gpurun_out/r2_tile8_$v.json 2> gpurun_out/r2_tile8_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_tile8_$v.json').read().strip().splitlines()[-1]); print('$v', 'N=8 C1 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
  done
done
