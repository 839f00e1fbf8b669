#!/bin/bash
# the round's closing check on one B200: GPU test suite, smoke(), the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log; tail -4 gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/f_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_pcie'), d['e2e']['device_resident']['value'], d['cpu_baseline']['value'], d['clocks'])
print({k: round(v['frac'], 3) for k, v in d.get('roofline_stress', {}).get('cases', {}).items()})
PY
