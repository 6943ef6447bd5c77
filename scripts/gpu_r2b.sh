#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_sca.py > gpurun_out/r2b_diag.log 2>&1
python -m pytest "tests/test_msda_gpu.py::test_full_size_backward_vs_oracle" tests/test_reference_gpu.py::test_bevrender_model_with_dropins_vs_reference_on_box -m gpu -q --tb=short -s > gpurun_out/r2b_tests.log 2>&1
tail -60 gpurun_out/r2b_diag.log; tail -30 gpurun_out/r2b_tests.log
