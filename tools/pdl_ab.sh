timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  for w in c4 c2 c3; do
    MDR_NO_PDL_UNUSED=$v timeout 200 python bench.py --workload $w --steps 600 --warmup 50 --no-cpu-baseline > gpurun_out/pdl${v}_${w}.json 2> gpurun_out/pdl${v}_${w}.err
  done
done
python - <<'PY'
import json
for v in ("0","1"):
    out=["MDR_NO_PDL="+v]
    for w in ("c4","c2","c3"):
        try:
            d=json.load(open("gpurun_out/pdl%s_%s.json"%(v,w))); out.append("%s %.1fus %.3f"%(w,d["roofline"]["launch_us"],d["roofline"]["frac"]))
        except Exception as ex: out.append(w+" ERR")
    print("  ".join(out))
PY
