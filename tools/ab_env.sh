# A/B of run-time switches of the in-tree library (same binary): usage  ab_env.sh "VAR=a" "VAR=b" ...
for e in "$@"; do
  echo "== $e"
  for w in ${WORKLOADS:-c3 c4}; do
    R=""; [ $w = c3 ] && R="--realizations 4000"; [ $w = c4 ] && R="--realizations 1024"; [ $w = c5 ] && R="--realizations 1024"
    env $e python bench.py --workload $w $R --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['value'], d['roofline']['frac'])"
  done
done
