mkdir -p gpurun_out/r2v
python -m pytest tests/test_grank_gpu.py -m gpu -x -q -k huge > gpurun_out/r2v/pytest2.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:merge_seq_kernel<.int.1024' --launch-skip 6 --launch-count 1 -o gpurun_out/r2v/seq_ba1m -f python tools/time_grank.py ba:1048576 8 1 > gpurun_out/r2v/ncu_ba.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:merge_seq_kernel<.int.1024' --launch-skip 6 --launch-count 1 -o gpurun_out/r2v/seq_r20 -f python tools/time_grank.py 20 8 1 > gpurun_out/r2v/ncu_r20.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:merge_dense_kernel<.int.4096' --launch-skip 4 --launch-count 1 -o gpurun_out/r2v/mid_r20 -f python tools/time_grank.py 20 8 1 > gpurun_out/r2v/ncu_mid.log 2>&1
