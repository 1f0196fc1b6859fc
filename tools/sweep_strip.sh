# sustained (power-capped) rate of the strip kernel variants: 12 back-to-back 4096-frame launches each, calls 4..12 averaged
for v in default g1 g4 w12 h20w20 h20w24 nounroll; do
  if [ $v = default ]; then unset B200GEO_LIB; else export B200GEO_LIB=$PWD/gpurun_variants/lib_$v.so; fi
  python tools/loop_strip.py 4096 2>&1 | python -c "
import sys
for l in sys.stdin:
    if 'per-call' in l:
        t=[float(x) for x in l.split('per-call ms:')[1].split()]
        print('$v', 'first %.2f' % t[0], 'sustained %.2f ms = %.0f Gpix/s' % (sum(t[3:])/len(t[3:]), 4096*1080*1920/(sum(t[3:])/len(t[3:]))/1e6))
"
  sleep 3
done
