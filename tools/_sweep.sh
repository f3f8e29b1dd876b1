python -m pytest tests/test_gpu_fused_warp.py -x -q -m gpu -k "slabs" 2>&1 | tail -3
timeout 400 python bench.py --loss ce --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ce_slab2.json 2> gpurun_out/r2_ce_slab2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_ce_slab2.json").read().strip().splitlines() if l.startswith('{')][-1])
print(round(d["value"]), round(d["ms_per_step"],3), {k.replace('arx_',''):(round(v["ms_per_step"],3), v["launches_per_step"]) for k,v in d["per_kernel"].items() if v["ms_per_step"]>0.2})
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pool_bwd_apply_slab -s 8 -c 2 python bench.py --loss ce --steps 1 --warmup 1 --no-cpu-baseline --no-graph 2>&1 | grep -E "gpu__time|dram__bytes|hit_rate"
