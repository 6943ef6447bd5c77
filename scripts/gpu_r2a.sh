#!/bin/bash
# round 2, first GPU pass: the reference-on-box parity tests + the new full-size backward tests
mkdir -p gpurun_out
python -m pytest tests/test_reference_gpu.py "tests/test_msda_gpu.py::test_full_size_backward_vs_oracle" \
  "tests/test_rpe_attn_gpu.py::test_rpe_attention_n100_backward_vs_dense_oracle_on_device" -m gpu -q --tb=short \
  > gpurun_out/r2a_tests.log 2>&1
tail -40 gpurun_out/r2a_tests.log
