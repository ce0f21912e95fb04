cd $GRAFT_REPO_ROOT
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | tail -n 6 > gpurun_out/r41_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r41_smoke.log 2>&1; echo smoke_rc=$? >> gpurun_out/r41_smoke.log
timeout 500 python bench.py > gpurun_out/r41_bench.json 2> gpurun_out/r41_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r41_ref.json 2> gpurun_out/r41_ref.err
echo done
