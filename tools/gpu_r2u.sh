#!/bin/bash
# compute-sanitizer on the round-2 kernels: racecheck (shared-memory hazards of the resident-column Jacobi, K2b, K3) and memcheck, 2 pixels
mkdir -p gpurun_out
for n in 12 16; do
SAN_N=$n timeout -s KILL 420 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python tools/sanitize_case.py > gpurun_out/r2u_race_$n.log 2>&1; echo "racecheck N=$n rc=$?"
grep -c "Race reported\|hazard" gpurun_out/r2u_race_$n.log; grep "RACECHECK SUMMARY\|^rc \|Error" gpurun_out/r2u_race_$n.log | head -5
done
SAN_N=12 timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_case.py > gpurun_out/r2u_mem_12.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|^rc " gpurun_out/r2u_mem_12.log | head
grep -B2 -A12 "Race reported" gpurun_out/r2u_race_12.log | head -60 | cut -c1-220
