# round 2, call lk4 (1 GPU): long keys only (empty device side)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "longer_than or commit_replaces or tokeniser" > gpurun_out/r02_lk4_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02_lk4_pytest.log | cut -c1-300
