python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_gputest_final.log; cat gpurun_out/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/r2_final_bench2.json 2> gpurun_out/r2_final_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_bench2.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],3))
PY
