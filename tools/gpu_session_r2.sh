# Round-2 measurement session on the GPU box (one B200): everything profiles/ holds is produced here.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_r2.sh'      then, here:  cp gpurun_out/profiles_r2/* profiles/
set -x
mkdir -p gpurun_out gpurun_out/profiles_r2
# 1. every launch of ONE tc32 step with its device time and DRAM bytes (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list_tc32_raw.csv python tools/one_step.py tc32 > gpurun_out/r2_one_step.log 2>&1
# 2. CUDA-event time of every C-ABI call of one step (not under ncu)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stock --profile-out gpurun_out/r2_per_call_tc32.jsonl > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
# 3. the dominant kernel, full capture
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_tc32 -s 3 -c 1 -f -o gpurun_out/r2_conv_tc32_full python tools/prof_conv.py --tc32 > gpurun_out/r2_prof_conv.log 2>&1
# 4. summaries (profiles/r2_dram_traffic.json is what bench.py reports as roofline.traffic), then the bench line itself
python tools/summarize_r2.py > gpurun_out/r2_summarize.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; tail -c 300 gpurun_out/r2_bench_line.err
timeout 400 python bench.py --workload viper --steps 10 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/r2_bench_viper.json 2> gpurun_out/r2_bench_viper.err; tail -c 300 gpurun_out/r2_bench_viper.err
timeout 200 python tools/microbench_flow.py > gpurun_out/r2_microbench_flow.json 2> gpurun_out/r2_microbench_flow.err
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_gpu_conv_tc32.py -q -x -k "case0 or case3 or case5 or residual or deconv or thin" > gpurun_out/r2_sanitizer_pytest.log 2>&1; tail -3 gpurun_out/r2_sanitizer_pytest.log; tail -5 gpurun_out/r2_sanitizer_memcheck.log
python tools/summarize_r2.py > gpurun_out/r2_summarize.log 2>&1
cp profiles/r2_* gpurun_out/profiles_r2/ 2>/dev/null
rm -f gpurun_out/r2_conv_tc32_full.ncu-rep.tmp
ls -la gpurun_out gpurun_out/profiles_r2 | tail -40
