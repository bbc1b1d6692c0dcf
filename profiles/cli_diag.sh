cd $GRAFT_REPO_ROOT
nvidia-smi -q | grep -i -m2 "persistence"
for i in 1 2 3; do python - <<'PY'
import time, sys
sys.path.insert(0,'.')
import pomfret_b200 as pb
g = pb.load_gpu()
t=time.time(); ctx = g.init([0]); t1=time.time(); b = g.batch_begin(ctx); t2=time.time()
print("init %.2fs batch_begin %.2fs" % (t1-t, t2-t1))
PY
done
python - <<'PY'
import sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench, conftest
conftest.run_synth('/tmp/s60', bench.synth_args(12, 2.5, 60, 100))
PY
for t in 1 4 16; do echo "== device ingest -t $t"; ./pomfret_b200/bin/pomfret methphase -t $t -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "engine ready|worker 0\]|windows done|used:"; done
echo "== host inflate -t 16"; POMFRET_HOST_INFLATE=1 ./pomfret_b200/bin/pomfret methphase -t 16 -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "engine ready|worker 0\]|windows done|used:"
echo "== strace-ish: time of first cuda calls under load"; for t in 16; do ./pomfret_b200/bin/pomfret methphase -t $t -c 60 -o /tmp/o --vcf /tmp/s60.vcf.gz /tmp/s60.bam 2>&1 | grep -E "^\[T::"; done
