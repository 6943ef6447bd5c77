#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_msda_gpu.py -m gpu -q -k host_pipeline 2>&1 | tail -1
run() { env $2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'e2e ms %.2f  %.4g q/s  %.1f GB/s' % (d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['aggregate_host_device_GBps']))"; }
run overlapped "A=1" | tee -a gpurun_out/e2e.txt
run joined "BEVR_E2E_JOIN_EVERY_STEP=1" | tee -a gpurun_out/e2e.txt
run overlapped_q2 "BEVR_QSPLIT=2" | tee -a gpurun_out/e2e.txt
run overlapped_q8 "BEVR_QSPLIT=8" | tee -a gpurun_out/e2e.txt
run overlapped_s3 "BEVR_NSLOTS=3" | tee -a gpurun_out/e2e.txt
run overlapped_bf16 "A=1" "--dtype bf16" | tee -a gpurun_out/e2e.txt
