python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --no-cpu --no-e2e --steps 5 $EXTRA 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms', d['roofline']['kernel_ms'], d['kernel_path'], d['checks']['image_sum'])"; }
run MXB_JIT_PIPE=0
run MXB_JIT_PIPE=1
run MXB_JIT_PREFETCH=1
run MXB_JIT_THREADS=768
run MXB_JIT_THREADS=704
run MXB_JIT_THREADS=576
run MXB_JIT=0
