# A/B: persisting-L2 window on/off (MDR_L2_PERSIST) for the three bench workloads
for v in 0 1; do
  for w in c4 c2 c3; do
    MDR_L2_PERSIST=$v python bench.py --workload $w --steps 600 --warmup 50 --no-cpu-baseline > gpurun_out/l2p${v}_${w}.json 2> gpurun_out/l2p${v}_${w}.err
  done
done
python - <<'PY'
import json
for v in ("0","1"):
    out=["l2_persist="+v]
    for w in ("c4","c2","c3"):
        try:
            d=json.load(open("gpurun_out/l2p%s_%s.json"%(v,w))); out.append("%s %.1fus %.3f"%(w,d["roofline"]["launch_us"],d["roofline"]["frac"]))
        except Exception as ex: out.append(w+" ERR")
    print("  ".join(out))
PY
