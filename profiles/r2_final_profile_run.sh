set -x
bash profiles/variant_sweep.sh "default nomask" "c4 c2 lecture5_1080 chess1080" > gpurun_out/r2_final_mask_ab.log 2>&1
for w in c1 c2 c3 c4 lecture5_1080; do
  ncu --set full --import-source on --clock-control none -k regex:render_frame --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_final_${w} python profiles/prof_one.py $w 3 > gpurun_out/r2_final_ncu_${w}.log 2>&1
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scaling-target > gpurun_out/r2_final_launch_plain.json 2> gpurun_out/r2_final_launch_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_c1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scaling-target > gpurun_out/r2_final_launch_ncu.json 2> gpurun_out/r2_final_launch_ncu.err
tail -3 gpurun_out/r2_final_mask_ab.log
