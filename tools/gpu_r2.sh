# Round-2 GPU job: tests, smoke, bench, profiles.  Usage (on the GPU box, via gpurun): bash tools/gpu_r2.sh TAG [steps...]
# steps: test smoke bench ref ncu_full launches step convs ew enc configs
set -x
TAG=${1:-r2}; shift
STEPS="${@:-test smoke bench}"
mkdir -p gpurun_out
for s in $STEPS; do
case $s in
test)   timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1; tail -15 gpurun_out/${TAG}_pytest.txt ;;
smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; tail -3 gpurun_out/${TAG}_smoke.txt ;;
bench)  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err ;;
ref)    timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 400 gpurun_out/${TAG}_bench_ref.json ;;
# ncu captures: gpurun_out/ is merged back only below 64 MiB, so the raw-page CSV is exported on the box and the
# .ncu-rep is kept only when small (a handful of kernels)
ncu_full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc' -o gpurun_out/${TAG}_prof_dom -f env CONV_ITERS=1 python tools/run_kernels_for_ncu.py > gpurun_out/${TAG}_ncu_dom.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_dom.log
        ncu -i gpurun_out/${TAG}_prof_dom.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_dom_raw.csv 2>/dev/null
        [ $(stat -c %s gpurun_out/${TAG}_prof_dom.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/${TAG}_prof_dom.ncu-rep ;;
ncu_stream) timeout 900 ncu --set full --clock-control none -k regex:'gn_stream|dwconv_tma|bicubic_tma|qkmax_tc' -o gpurun_out/${TAG}_prof_ew -f env CONV_ITERS=0 python tools/run_kernels_for_ncu.py > gpurun_out/${TAG}_ncu_ew.log 2>&1
        ncu -i gpurun_out/${TAG}_prof_ew.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_ew_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_prof_ew.ncu-rep ;;
ncu_dw) timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv_tma' -c 2 -o gpurun_out/${TAG}_prof_dw -f env CONV_ITERS=0 python tools/run_kernels_for_ncu.py > gpurun_out/${TAG}_ncu_dw.log 2>&1
        ncu -i gpurun_out/${TAG}_prof_dw.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_dw_raw.csv 2>/dev/null ;;
ncu_enc) ENC_STAGES=${ENC_STAGES:-2} timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc|gemm_tc' -o gpurun_out/${TAG}_prof_enc -f python tools/run_encoder_for_ncu.py > gpurun_out/${TAG}_ncu_enc.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_enc.log
        ncu -i gpurun_out/${TAG}_prof_enc.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_enc_raw.csv 2>/dev/null
        [ $(stat -c %s gpurun_out/${TAG}_prof_enc.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/${TAG}_prof_enc.ncu-rep ;;
enc)    ENC_TIME=1 timeout 300 python tools/run_encoder_for_ncu.py > gpurun_out/${TAG}_encoder_gemms.txt 2>&1; tail -40 gpurun_out/${TAG}_encoder_gemms.txt
        echo "---- CAMRADEPTH_TC_PIPE=0" >> gpurun_out/${TAG}_encoder_gemms.txt
        CAMRADEPTH_TC_PIPE=0 ENC_TIME=1 ENC_STAGES=1,2 timeout 300 python tools/run_encoder_for_ncu.py >> gpurun_out/${TAG}_encoder_gemms.txt 2>&1; tail -20 gpurun_out/${TAG}_encoder_gemms.txt ;;
launches) CAMRADEPTH_PROFILE_TIMED=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --quick --no-graph > gpurun_out/${TAG}_launches.log 2>&1
        python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv 70 > gpurun_out/${TAG}_launches_summary.txt 2>&1; head -30 gpurun_out/${TAG}_launches_summary.txt ;;
step)   timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/${TAG}_step_profile.txt 2>&1; head -40 gpurun_out/${TAG}_step_profile.txt ;;
convs)  timeout 300 python tools/profile_convs.py 32 > gpurun_out/${TAG}_conv_profile.txt 2>&1; tail -30 gpurun_out/${TAG}_conv_profile.txt ;;
ew)     timeout 300 python tools/bench_elementwise.py > gpurun_out/${TAG}_elementwise.txt 2>&1 ;;
configs) timeout 900 python tools/bench_configs.py > gpurun_out/${TAG}_bench_configs.jsonl 2> gpurun_out/${TAG}_bench_configs.err; cat gpurun_out/${TAG}_bench_configs.jsonl ;;
ab)     # A/B of engine switches: AB="CAMRADEPTH_SPLIT=;CAMRADEPTH_LEAF_STAGES=" -> one quick bench line per setting
        echo "${AB}" | tr ';' '\n' | while read -r kv; do
          echo "== ${kv:-default}" >> gpurun_out/${TAG}_ab.txt
          env ${kv} timeout 600 python bench.py --quick --steps 10 --warmup 3 >> gpurun_out/${TAG}_ab.txt 2>/dev/null
          [ -n "${AB_LAT}" ] && env ${kv} timeout 300 python tools/bench_latency.py >> gpurun_out/${TAG}_ab.txt 2>/dev/null
        done; cat gpurun_out/${TAG}_ab.txt ;;
det)    timeout 300 python tools/det_check.py ${DET_VARIANT:-base} > gpurun_out/${TAG}_det_check.txt 2>&1; tail -20 gpurun_out/${TAG}_det_check.txt ;;
dp2)    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py supervised_seg fp32 > gpurun_out/${TAG}_dp_check.txt 2>&1; tail -4 gpurun_out/${TAG}_dp_check.txt
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py base bf16 >> gpurun_out/${TAG}_dp_check.txt 2>&1; tail -3 gpurun_out/${TAG}_dp_check.txt ;;
ncu_k)  # full-set ncu capture of the kernels matching NCU_K (regex) inside one eager training step
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K}" -c ${NCU_C:-9} -o gpurun_out/${TAG}_prof_k -f python bench.py --steps 1 --warmup 1 --quick --no-graph > gpurun_out/${TAG}_ncu_k.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_k.log
        ncu -i gpurun_out/${TAG}_prof_k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_k_raw.csv 2>/dev/null
        ncu -i gpurun_out/${TAG}_prof_k.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_k_details.csv 2>/dev/null
        [ $(stat -c %s gpurun_out/${TAG}_prof_k.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/${TAG}_prof_k.ncu-rep ;;
ablib)  # same-box A/B of library builds: ABLIBS="ab_libs/lib_head.so ab_libs/lib_lane0.so default"
        for L in ${ABLIBS}; do
          echo "==== ${L}" >> gpurun_out/${TAG}_ablib.txt
          if [ "$L" = default ]; then unset CAMRADEPTH_LIB; else export CAMRADEPTH_LIB=$PWD/$L; fi
          timeout 300 python tools/profile_convs.py 32 2>&1 | grep "depth_upsample.4\|conv launches total" >> gpurun_out/${TAG}_ablib.txt
          timeout 600 python bench.py --quick --steps 10 --warmup 3 2>/dev/null | cut -c1-200 >> gpurun_out/${TAG}_ablib.txt
        done; unset CAMRADEPTH_LIB; cat gpurun_out/${TAG}_ablib.txt ;;
dpo)    N=${NGPU:-2}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dp_overhead.py 30 > gpurun_out/${TAG}_dp_overhead_${N}gpu.txt 2>&1; grep "ms/step" gpurun_out/${TAG}_dp_overhead_${N}gpu.txt || tail -20 gpurun_out/${TAG}_dp_overhead_${N}gpu.txt ;;
benchN) N=${NGPU:-2}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; tail -c 1500 gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err ;;
esac
done
du -sh gpurun_out; ls -la gpurun_out | tail -20
