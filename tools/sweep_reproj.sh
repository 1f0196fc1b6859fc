for v in g1t256c2 g2t256c1 g2t256c2 g2t192c2 g2t128c3; do
  echo "== $v"
  B200GEO_LIB=$PWD/gpurun_variants/lib_$v.so python tools/bench_score.py 200000 2>&1 | grep "reproj    exact=False"
  B200GEO_LIB=$PWD/gpurun_variants/lib_$v.so python tools/fuzz_score.py 150 9 2>&1 | tail -1
done
