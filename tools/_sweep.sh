python -m pytest tests/test_gpu_fused_warp.py tests/test_gpu_parity.py -x -q -m gpu -k "slabs or deterministic" 2>&1 | tail -15
for sl in 1 0; do
  ARX_CATALOG_SLABS=$sl timeout 400 python bench.py --loss ce --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ce_slab$sl.json 2> gpurun_out/r2_ce_slab$sl.err
  python - $sl <<PY
import json,sys
sl=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r2_ce_slab%s.json"%sl).read().strip().splitlines() if l.startswith('{')][-1])
    print('slabs', sl, round(d["value"]), round(d["ms_per_step"],3), 'loss', d["e2e"].get("last_loss"))
    print({k.replace('arx_',''):(round(v["ms_per_step"],3), v["launches_per_step"]) for k,v in d["per_kernel"].items() if v["ms_per_step"]>0.2})
except Exception as e:
    print('FAILED', e); print(open("gpurun_out/r2_ce_slab%s.err"%sl).read()[-1500:])
PY
done
