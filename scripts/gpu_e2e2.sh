#!/bin/bash
run() { env $2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'e2e ms %.2f  %.4g q/s  %.1f GB/s' % (d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['aggregate_host_device_GBps']))"; }
for i in 1 2; do
run overlapped_q1 "BEVR_QSPLIT=1" | tee -a gpurun_out/e2e.txt
run overlapped_q2 "BEVR_QSPLIT=2" | tee -a gpurun_out/e2e.txt
run overlapped_q4 "BEVR_QSPLIT=4" | tee -a gpurun_out/e2e.txt
run overlapped_q2s3 "BEVR_QSPLIT=2 BEVR_NSLOTS=3" | tee -a gpurun_out/e2e.txt
run joined_q2 "BEVR_QSPLIT=2 BEVR_E2E_JOIN_EVERY_STEP=1" | tee -a gpurun_out/e2e.txt
done
