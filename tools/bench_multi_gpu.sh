# usage: bash tools/bench_multi_gpu.sh N "1 0"   -> bench.py on N GPUs with the peer-memory exchange (1) and / or the NCCL collectives (0); one-line summaries
n=${1:-2}
for peer in ${2:-1}; do
  ARX_PEER=$peer timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$peer bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mg_n${n}_peer$peer.json 2> gpurun_out/r2_mg_n${n}_peer$peer.err
  echo "rc=$? peer=$peer"
  python - $n $peer <<PY
import json,sys
n,peer=sys.argv[1:3]
try:
    d=json.loads([l for l in open("gpurun_out/r2_mg_n%s_peer%s.json"%(n,peer)).read().strip().splitlines() if l.startswith('{')][-1])
    print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],4), 'e2e', round(d["e2e"]["value"]), d.get("sharded_check"))
    print({k.replace('arx_',''):round(v["avg_us"],1) for k,v in d["per_kernel"].items()})
except Exception as e:
    print('FAILED', e); print(open("gpurun_out/r2_mg_n%s_peer%s.err"%(n,peer)).read()[-1500:])
PY
done
