mkdir -p gpurun_out
echo "== default direct w20n2" > gpurun_out/run4.log
timeout 300 python tools/check_strip.py 256 >> gpurun_out/run4.log 2>&1
for v in d24 d16 d20n3 d24nu; do
  echo "== $v" >> gpurun_out/run4.log
  B200GEO_LIB=$PWD/gpurun_variants/lib_$v.so timeout 300 python tools/check_strip.py 256 2>&1 | grep "variant 0" >> gpurun_out/run4.log
done
timeout 600 python tools/fuzz_warp.py 200 8 >> gpurun_out/run4.log 2>&1
cat gpurun_out/run4.log
