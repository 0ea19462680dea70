# round 2, call lk3 (1 GPU): whole GPU suite after the long-key store (tests that expected MRHBM_E_KEY follow the new semantics)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_lk3_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02_lk3_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2 | cut -c1-200
