run() {
  env "$@" python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2_sw.json 2> gpurun_out/r2_sw.err
  python - "$*" <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/r2_sw.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],4), {k.replace('arx_',''):round(v["avg_us"],1) for k,v in d["per_kernel"].items() if "apply" in k or "plan" in k or "fwd_many" in k})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open("gpurun_out/r2_sw.err").read()[-800:])
PY
}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_shapes.py -x -q -m gpu 2>&1 | tail -4
run A=1
run ARX_TUNE=apply_flat=1
run ARX_TUNE=apply_ctas_per_sm=2
run ARX_TUNE=apply_ctas_per_sm=6
run ARX_TUNE=apply_contig=1
