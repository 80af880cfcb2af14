tools/run_env_variants.sh base "FMHR_COV_SMEM=24576" "FMHR_COV_SMEM=20480" "FMHR_COV_SMEM=34816" > gpurun_out/s3f_env.txt 2>&1; cat gpurun_out/s3f_env.txt
FMHR_COV_SMEM=24576 FMHR_B200_LIB=/root/repo/variants/libfmhr_trace.so python tools/trace_timeline.py > gpurun_out/s3f_trace_cov4.txt 2>&1; tail -15 gpurun_out/s3f_trace_cov4.txt
