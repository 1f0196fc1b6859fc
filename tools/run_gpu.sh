mkdir -p gpurun_out
export B200GEO_LIB=$PWD/gpurun_variants/lib_w16n1.so
python tools/prof_strip.py 64 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_strip_kernel -s 3 -c 1 -f -o gpurun_out/r02_strip_v2 python tools/prof_strip.py 64 > gpurun_out/ncu.log 2>&1
python tools/prof_strip.py 64 nn > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_strip_kernel -s 3 -c 1 -f -o gpurun_out/r02_strip_v2_nn python tools/prof_strip.py 64 nn > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu.log gpurun_out/ncu2.log
