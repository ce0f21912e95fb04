cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 1 --warmup 3 --no-attribution --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r28_plain.json 2> gpurun_out/r28_plain.err || { echo plain failed; exit 1; }
CTC_BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r28_launches.csv $CMD > gpurun_out/r28_ncu.log 2>&1
CTC_BENCH_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -c 16 -f -o gpurun_out/r28_gemm $CMD > gpurun_out/r28_ncu2.log 2>&1
CTC_BENCH_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_|peg_tma|patchify|latent_proj_mma" -c 12 -f -o gpurun_out/r28_misc $CMD > gpurun_out/r28_ncu3.log 2>&1
echo done
