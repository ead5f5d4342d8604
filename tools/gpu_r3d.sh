#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python tools/parity_large.py 2048 > gpurun_out/r3d_parity_large.json 2> gpurun_out/r3d_parity_large.err; echo "rc=$?"; tail -3 gpurun_out/r3d_parity_large.err | cut -c1-200; cat gpurun_out/r3d_parity_large.json | tr '\n' ' ' | cut -c1-900
