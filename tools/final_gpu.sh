#!/bin/bash
# Round-end GPU pass (one GPU): tests, smoke, both bench arms, the ncu launch list of the bench command, DRAM traffic.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 300 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_dram_bytes.csv python tools/k1_prof.py cfg2_150bp 1000000 > gpurun_out/r02_ncu_dram.log 2>&1; echo "ncu dram rc=$?"
