#!/bin/bash
# round-2 GPU session 3: new parity tests (adv random, adver 3, update-scale), ml-1m shape, Video APR-phase KAT
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_parity.log
tail -25 gpurun_out/r2c_pytest_parity.log
timeout 1500 python -m pytest tests/test_gpu_ml1m_shape.py -m gpu -x -q -s > gpurun_out/r2c_pytest_ml1m.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_ml1m.log
tail -25 gpurun_out/r2c_pytest_ml1m.log
timeout 1500 python -m pytest tests/test_gpu_video_apr_kat.py -m gpu -x -q -s > gpurun_out/r2c_pytest_video.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_video.log
tail -30 gpurun_out/r2c_pytest_video.log
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_recommender.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r2c_pytest_rest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_rest.log
tail -8 gpurun_out/r2c_pytest_rest.log
