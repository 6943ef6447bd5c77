#!/bin/bash
timeout 600 python -m pytest tests/test_msda_gpu.py -m gpu -q -x 2>&1 | tail -1
run() { BEVR_SIGMA_PX=$4 BEVR_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 sigma=$4', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" "" 2.0
run default "" "" 1.0
run o3_64regs_3ctas scripts/_build/libbevr_o3.so "" 1.0
run o4_64regs_4ctas scripts/_build/libbevr_o4.so "" 1.0
run r75_maxbins512 scripts/_build/libbevr_r75m512.so "" 1.0
