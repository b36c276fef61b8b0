set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2_per_call_tc32.jsonl > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; tail -c 300 gpurun_out/r2_bench_line.err
timeout 600 python bench.py --workload viper --steps 10 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench_viper.json 2> gpurun_out/r2_bench_viper.err; tail -c 300 gpurun_out/r2_bench_viper.err
timeout 300 python tools/microbench_flow.py > gpurun_out/r2_microbench_flow.json 2> gpurun_out/r2_microbench_flow.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list_tc32_raw.csv python tools/one_step.py tc32 > gpurun_out/r2_one_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_tc32 -s 3 -c 1 -f -o gpurun_out/r2_conv_tc32_full python tools/prof_conv.py --tc32 > gpurun_out/r2_prof_conv.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_gpu_conv_tc32.py -q -x -k "case0 or case3 or case5 or residual or deconv" > gpurun_out/r2_sanitizer_pytest.log 2>&1; tail -3 gpurun_out/r2_sanitizer_pytest.log; tail -5 gpurun_out/r2_sanitizer_memcheck.log
ls -la gpurun_out | tail -20
