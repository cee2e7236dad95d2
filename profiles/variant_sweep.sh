# usage: profiles/variant_sweep.sh "<variants>" "<workloads>"   (tuning aid: kernel ms per frame of build_variants/<v>; "default" = in-tree build)
for v in ${1:-default}; do
  for w in ${2:-c4 lecture5_1080}; do
    if [ "$v" = default ]; then unset C2RT_LIB_DIR; else export C2RT_LIB_DIR=build_variants/$v; fi
    echo -n "$v: "; python profiles/prof_one.py $w 6 2>&1 | tail -1
  done
done
