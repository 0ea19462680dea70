mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -8 gpurun_out/pytest.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value']/1e9,'Gpairs/s', d['ms_per_step'], d['roofline']['stages_ms'], d['e2e'], d['config'])"; tail -3 gpurun_out/bench.err
