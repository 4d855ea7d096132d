mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solves.py tests/test_gpu_properties.py -m gpu -x -q -k "nlpoisson or newton or nonlinear or quad_models" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c2_pytest.log
timeout 300 tools/ab_libs.sh "python tools/probe.py nlpoisson 4096 20 gather,atomic" c4sym.so cur.so > gpurun_out/c2_ab_c4.log 2>&1
grep -E "===|best" gpurun_out/c2_ab_c4.log
