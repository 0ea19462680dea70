# round 2, call heap (1 GPU): the iterator merges a partition's runs through a binary heap instead of a linear scan: GPU suite + smoke
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r02_heap_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02_heap_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2 | cut -c1-200
