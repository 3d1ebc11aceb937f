# A/B of differently built libraries: tools/run_ab.sh name1=path1 name2=path2 ... (path "-" = in-tree build)
for spec in "$@"; do
  name=${spec%%=*}; path=${spec#*=}
  for w in c4 c2 c3; do
    if [ "$path" = "-" ]; then unset MDR_LIB_PATH; else export MDR_LIB_PATH=$PWD/$path; fi
    python bench.py --workload $w --steps 600 --warmup 50 --no-cpu-baseline > gpurun_out/ab_${name}_${w}.json 2> gpurun_out/ab_${name}_${w}.err
  done
done
unset MDR_LIB_PATH
python - "$@" <<'PY'
import json, sys
for spec in sys.argv[1:]:
    name=spec.split("=")[0]; out=[name]
    for w in ("c4","c2","c3"):
        try:
            d=json.load(open("gpurun_out/ab_%s_%s.json"%(name,w))); out.append("%s %.1fus %.3f"%(w,d["roofline"]["launch_us"],d["roofline"]["frac"]))
        except Exception as ex: out.append(w+" ERR")
    print("  ".join(out))
PY
