cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r38_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 77 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r38_memcheck_smoke.log 2>&1
echo smoke_rc=$?
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_dropin.py "tests/test_gpu_ops.py::test_attention_fwd_tcgen05_matches_reference_and_mma_sync" "tests/test_gpu_ops.py::test_attention_bwd_dq_on_tcgen05" tests/test_gpu_ops.py::test_peg_frames_matches_dense_peg tests/test_gpu_ops.py::test_latent_proj_sim -q -m gpu -x -k "not 24-24 and not 294912" > gpurun_out/r38_memcheck_tests.log 2>&1
echo tests_rc=$?
