# A/B harness used while tuning (variants are built into build/ with tools/build_variant.sh; see profiles/*_notes.md)
for lib in build/lib_*.so; do
  echo "== $lib"
  for w in ${WORKLOADS:-c3 c4}; do
    R=""; [ $w = c3 ] && R="--realizations 4000"; [ $w = c4 ] && R="--realizations 1024"; [ $w = c5 ] && R="--realizations 1024"
    ONEKA_B200_LIB=$PWD/$lib python bench.py --workload $w $R --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['value'], d['roofline']['frac'])"
  done
done
