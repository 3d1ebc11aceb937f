# c3big (1000 x 1000 houses) through the split pipelined kernel with forced cluster sizes; prints the co-resident clusters
tag=${1:-x}
for k in 0 2 4 5 8; do
  MDR_VERBOSE=1 MDR_SPLIT_K=$k timeout 120 python bench.py --workload c3big --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/c3big_k${k}_$tag.json 2> gpurun_out/c3big_k${k}_$tag.err
  grep "\[mdr\]" gpurun_out/c3big_k${k}_$tag.err | head -1
  python -c "
import json;d=json.load(open('gpurun_out/c3big_k${k}_$tag.json'));print('k=$k', 'us/step %.2f'%(d['ms_per_step']*1e3), d['launch'])"
done
