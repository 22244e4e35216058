#!/bin/bash
# round-end single-GPU measurements; everything lands in gpurun_out/
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r1_rmat16.json 2> gpurun_out/err1.log; python scratch/show.py gpurun_out/bench_r1_rmat16.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_r1_ref.json 2> gpurun_out/err2.log; cat gpurun_out/bench_r1_ref.json | cut -c1-400
timeout 300 python bench.py --workload rmat20mc --steps 3 --warmup 3 > gpurun_out/bench_r1_rmat20mc.json 2> gpurun_out/err3.log; python scratch/show.py gpurun_out/bench_r1_rmat20mc.json
timeout 400 python bench.py --workload rmat22 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_rmat22.json 2> gpurun_out/err4.log; python scratch/show.py gpurun_out/bench_r1_rmat22.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_bench_rmat16.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:merge_ -s 612 -c 12 -o gpurun_out/prof_r1_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
