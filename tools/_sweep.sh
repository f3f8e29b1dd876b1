python -m pytest tests/test_gpu_token_pool.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_t5.log
python -m pytest tests/test_ref_golden.py -x -q -m gpu -k "nondefault" 2>&1 | tail -8 >> gpurun_out/r2_t5.log
cat gpurun_out/r2_t5.log
for cfg in "heavy=32,apply_flat=0" "heavy=16,apply_flat=0" "heavy=16,apply_flat=1" "heavy=8,apply_flat=0" "heavy=8,apply_flat=1"; do
  ARX_TUNE=$cfg python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2_sw.json 2> gpurun_out/r2_sw.err
  python - "$cfg" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2_sw.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],4), {k:round(v["avg_us"],1) for k,v in d["per_kernel"].items() if "apply" in k or "plan" in k})
PY
done
