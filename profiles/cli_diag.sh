cd $GRAFT_REPO_ROOT
python - <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench, conftest
conftest.run_synth('/tmp/s60', bench.synth_args(16, 5, 60, 100))
PY
ls -la /tmp/s60.bam
for i in 1 2; do echo "== ours -t 16 (run $i)"; ./pomfret_b200/bin/pomfret methphase -t 16 -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "^\[T::|used:|inflated on"; done
echo "== ours POMFRET_WORKERS=6"; POMFRET_WORKERS=6 ./pomfret_b200/bin/pomfret methphase -t 16 -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "^\[T::|used:"
echo "== ours POMFRET_WORKERS=1"; POMFRET_WORKERS=1 ./pomfret_b200/bin/pomfret methphase -t 16 -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "^\[T::|used:"
echo "== ref -t 16"; ./oracle/_ref/pomfret methphase -t 16 -c 60 -o /tmp/r --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "used:"
