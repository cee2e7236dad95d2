# Final captures of round 2, second session (kernels of commit a222ebc and later): one ncu --set full report per workload and the
# launch list of the bench command.  Run under gpurun; the reports come back in gpurun_out/ and are summarised by
# profiles/ncu_summary.py / ncu_executed.py / ncu_lines.py into profiles/r2_s2_*.
set -x
for w in c1 c2 c3 c4 lecture5_1080; do
  ncu --set full --import-source on --clock-control none -k regex:render_frame --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_s2_${w} python profiles/prof_one.py $w 3 > gpurun_out/r2_s2_ncu_${w}.log 2>&1
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scaling-target > gpurun_out/r2_s2_launch_plain.json 2> gpurun_out/r2_s2_launch_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_s2_c1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scaling-target > gpurun_out/r2_s2_launch_ncu.json 2> gpurun_out/r2_s2_launch_ncu.err
ls -la gpurun_out | tail -12
