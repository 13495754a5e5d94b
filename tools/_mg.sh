cd /root/repo
for wl in c3 c5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --workload $wl $( [ $wl = c5 ] && echo "--kpts 4096" ) > gpurun_out/bench_2gpu_$wl.json 2> gpurun_out/bench_2gpu_$wl.err; cat gpurun_out/bench_2gpu_$wl.json | cut -c1-900; tail -2 gpurun_out/bench_2gpu_$wl.err
done
python bench.py --steps 10 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_1gpu_c3.json 2>/dev/null; cut -c1-400 gpurun_out/bench_1gpu_c3.json
