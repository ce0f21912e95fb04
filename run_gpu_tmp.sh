cd $GRAFT_REPO_ROOT
timeout 600 python tools/prof_step.py 8 > gpurun_out/r6_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"peg_kernel|patchify_ln_bwd|patchify_ln_fwd|attn_bwd_dkv_kernel<96|attn_fwd_kernel<192|gemm_tcgen05_kernel<256, 1>" \
   -c 14 -o gpurun_out/r6_prof python tools/prof_step.py 8 > gpurun_out/r6_ncu.log 2>&1
echo done
