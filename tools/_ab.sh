cd /root/repo
for nb in 0 1 0 1; do
LGB200_NUMA_BIND=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('bind=$nb', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('host_numa_bound'), d['clocks'])
"
done | tee gpurun_out/numa_ab.log
nvidia-smi topo -m 2>/dev/null | head -20 | tee -a gpurun_out/numa_ab.log; nproc | tee -a gpurun_out/numa_ab.log
