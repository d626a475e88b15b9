# Refresh of the measurement evidence under gpurun_out/ (copied into profiles/ afterwards).  FULL=1 adds the
# `ncu --set full` capture of the dominant kernels.
set -x
mkdir -p gpurun_out
if [ "${FULL:-0}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc|gn_stream|dwconv_tma|bicubic_tma|qkmax_tc' -o gpurun_out/prof_r1b -f env CONV_ITERS=1 python tools/run_kernels_for_ncu.py > gpurun_out/ncu_full_r1b.log 2>&1
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 0 --quick --no-graph > gpurun_out/launches_r1b.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_r1b.csv 60 last_step > gpurun_out/launches_r1b_summary.txt 2>&1
timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/step_profile_r1b.txt 2>&1
timeout 300 python tools/profile_convs.py 32 > gpurun_out/conv_profile_r1b.txt 2>&1
timeout 300 python tools/bench_elementwise.py > gpurun_out/elementwise_r1b.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1b.json 2> gpurun_out/bench_ref_r1b.err
tail -c 600 gpurun_out/bench_r1b.json; tail -c 300 gpurun_out/bench_ref_r1b.json; head -12 gpurun_out/launches_r1b_summary.txt
