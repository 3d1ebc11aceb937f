# compute-sanitizer evidence run (one GPU): memcheck, racecheck, synccheck and initcheck over the kernels of
# libmdr_b200.so only (--kernel-name kns=mdr), driven by tools/sanitize_driver.py.  Logs -> gpurun_out/sanitize_*_<tag>.log
# (copy the summaries into profiles/).  Usage: bash tools/sanitize.sh r02 [families]
tag=${1:-r02}
fam=${2:-pipe,generic,fused,cluster,populate}
for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --kernel-name kns=mdr --print-limit 20 --error-exitcode 7 \
        --log-file gpurun_out/sanitize_${tool}_$tag.log python tools/sanitize_driver.py $fam > gpurun_out/sanitize_${tool}_$tag.out 2>&1
    echo "$tool exit $?" >> gpurun_out/sanitize_${tool}_$tag.out
    tail -3 gpurun_out/sanitize_${tool}_$tag.log
done
