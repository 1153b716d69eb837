python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --no-cpu --no-e2e --steps 5 $EXTRA 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms', d['roofline']['kernel_ms'], d['kernel_path'], d['checks'])"; }
for EXTRA in "" "--no-image"; do
echo "#### EXTRA=$EXTRA"
run MXB_JIT_THREADS=512 MXB_JIT_MINBLOCKS=1
run MXB_JIT_THREADS=256 MXB_JIT_MINBLOCKS=2
run MXB_JIT_THREADS=640 MXB_JIT_MINBLOCKS=1
run MXB_JIT=0
done
