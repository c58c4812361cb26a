run() { PTAR_MIN_CTAS=$1 PTAR_NVCC_EXTRA="$2" python -c "import __graft_entry__ as g; g.build(force=True)" && python bench.py --steps 3 --warmup 3 --no-cpu $3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('CFG', '$1', '$2', '$3', round(d['value']), {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()}, round(d['roofline']['frac'],4))"; }
run 4 "" ""
run 4 "-DGEN_UNROLL=2" ""
run 3 "-DGEN_UNROLL=2" ""
run 4 "-DGEN_WHITE_FP32" ""
run 4 "-DPHILOX_ROUNDS=7" ""
run 4 "" "--chunk 512"
run 4 "-DGEN_WHITE_FP32" "--merged-white"
