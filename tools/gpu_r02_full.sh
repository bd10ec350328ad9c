# Round-2 full GPU validation + measurements (run under gpurun from the repo root).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm.py -q -x -p no:cacheprovider > gpurun_out/r02_pytest_gemm.log 2>&1; echo GEMM_RC=$?; tail -5 gpurun_out/r02_pytest_gemm.log
python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider --ignore tests/test_gpu_gemm.py > gpurun_out/r02_pytest.log 2>&1; tail -12 gpurun_out/r02_pytest.log
timeout 300 python tools/gemm_bench.py > gpurun_out/r02_gemm_bench.log 2>&1; tail -4 gpurun_out/r02_gemm_bench.log | cut -c1-420
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cat gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2>/dev/null; cat gpurun_out/r02_bench_reference.json
timeout 600 python tools/model_bench.py opt 2> gpurun_out/model_opt.err | grep '^{' | tail -1 > gpurun_out/r02_model_opt.json; cat gpurun_out/r02_model_opt.json; tail -2 gpurun_out/model_opt.err
timeout 600 python tools/model_bench.py resnet 2> gpurun_out/model_resnet.err | grep '^{' | tail -1 > gpurun_out/r02_model_resnet.json; cat gpurun_out/r02_model_resnet.json; tail -2 gpurun_out/model_resnet.err
timeout 600 python tools/model_bench.py bert 2> gpurun_out/model_bert.err | grep '^{' | tail -1 > gpurun_out/r02_model_bert.json; cat gpurun_out/r02_model_bert.json; tail -2 gpurun_out/model_bert.err
python tools/calib_bench.py 2>/dev/null | tail -1 > gpurun_out/r02_calibration.json; cat gpurun_out/r02_calibration.json
# DRAM traffic where it is observable: 16384^2 (1.07 GB per launch >> 126 MB L2), and the headline size for the read side
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:antq_stream -s 2 -c 2 --csv --log-file gpurun_out/r02_traffic_16384.csv python tools/quick_bench.py --rows 16384 --cols 16384 --nb 2 --reps 1 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:antq_stream -s 10 -c 4 --csv --log-file gpurun_out/r02_traffic_4096.csv python tools/quick_bench.py --reps 1 > /dev/null 2>&1
tail -3 gpurun_out/r02_traffic_16384.csv; tail -2 gpurun_out/r02_traffic_4096.csv
# launch list of the bench command (every launch with its device time)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
grep -c antq_stream gpurun_out/r02_launches.csv
# full captures: headline kernel, closed-form kernel (int-8), GEMM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:antq_stream_kernel -s 12 -c 1 -f -o gpurun_out/r02_stream python tools/quick_bench.py --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:antq_pu_stream -s 12 -c 1 -f -o gpurun_out/r02_pu_int8 python tools/quick_bench.py --kind int --bit 8 --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:antq_linear_p4 -s 2 -c 1 -f -o gpurun_out/r02_gemm python tools/gemm_probe.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/sweep.py --out gpurun_out/r02_sweep.jsonl > gpurun_out/r02_sweep.log 2>&1; wc -l gpurun_out/r02_sweep.jsonl
