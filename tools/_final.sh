set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_gputest_final.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
timeout 400 python bench.py --loss warp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_warp.json 2> gpurun_out/r2_final_warp.err
for w in c3 c5 lstm_cell; do python bench_extra.py --workload $w > gpurun_out/r2_final_$w.json 2> gpurun_out/r2_final_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2_final_nculist.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pool_fwd_flat_many|pool_bwd_apply_kernel|ce_kernel|mw_prep|mw_post|plan_count|plan_fill|plan_alloc" -s 28 -c 14 -o gpurun_out/r2_final_step python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2_final_ncufull.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pool_bwd_apply_slab" -s 8 -c 2 -o gpurun_out/r2_final_slab python bench.py --loss ce --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2_final_ncuslab.log 2>&1
python tools/trace_step.py > gpurun_out/r2_final_timeline.txt 2> gpurun_out/r2_final_timeline.err
tail -3 gpurun_out/r2_gputest_final.log; head -c 400 gpurun_out/r2_final_bench.json; ls -la gpurun_out | tail -20
