#!/bin/bash
# two devices: the in-library multi-device tests and the new cloud forms split over two devices
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -s KILL 600 python -m pytest tests/test_multidevice_gpu.py -m gpu -q --durations=3 > gpurun_out/r4d_tests.log 2>&1; echo "multidevice rc=$?"; tail -6 gpurun_out/r4d_tests.log | cut -c1-200
AB200_DEVICES=2 timeout -s KILL 600 python -m pytest tests/test_cloud_more_gpu.py tests/test_cloud_gpu.py -m gpu -q > gpurun_out/r4d_cloud2.log 2>&1; echo "cloud on 2 devices rc=$?"; tail -3 gpurun_out/r4d_cloud2.log | cut -c1-200
