set -e
cd $GRAFT_REPO_ROOT
mkdir -p /tmp/w && python - <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench, conftest
conftest.run_synth('/tmp/w/s60', bench.synth_args(12, 2.5, 60, 100))
PY
ls -la /tmp/w
for t in 16 6; do
  echo "== ours -t $t"
  ./pomfret_b200/bin/pomfret methphase -t $t -c 60 -o /tmp/w/o --vcf /tmp/w/s60.vcf.gz /tmp/w/s60.bam 2>&1 | grep -E "^\[T::|used:" 
done
echo "== ours -t 16 --windows-per-batch 2"
./pomfret_b200/bin/pomfret methphase -t 16 --windows-per-batch 2 -c 60 -o /tmp/w/o --vcf /tmp/w/s60.vcf.gz /tmp/w/s60.bam 2>&1 | grep -E "^\[T::|used:"
echo "== ref -t 16"
./oracle/_ref/pomfret methphase -t 16 -c 60 -o /tmp/w/r --vcf /tmp/w/s60.vcf.gz /tmp/w/s60.bam 2>&1 | grep -E "used:"
echo "== ours report"
./pomfret_b200/bin/pomfret report -t 16 -c 60 --chunk-size 50000 --chunk-stride 100000 -o /tmp/w/o --vcf /tmp/w/s60.vcf.gz /tmp/w/s60.bam 2>&1 | grep -E "^\[T::|used:|windows "
