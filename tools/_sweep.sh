python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py tests/test_ref_golden.py -x -q -m gpu -k "mw or glue or hmf" 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_sw.json 2> gpurun_out/r2_sw.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_sw.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), {k.replace('arx_',''):round(v["avg_us"],1) for k,v in d["per_kernel"].items()})
PY
