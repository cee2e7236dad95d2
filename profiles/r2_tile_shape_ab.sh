for v in default tile32x4; do
  if [ "$v" = default ]; then unset C2RT_LIB_DIR; else export C2RT_LIB_DIR=build_variants/$v; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_tile_$v.json 2> gpurun_out/r2_tile_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_tile_$v.json').read().strip().splitlines()[-1]); print('$v', 'N=2 C1', d['ms_per_step'], [ (t['workload'][:2], round(t['ms_per_step'],3), round(t['ms_per_step_1gpu_same_run'],3)) for t in d['scaling_targets']])"
  python profiles/prof_one.py c1 5 | tail -1
done
