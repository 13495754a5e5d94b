cd /root/repo
for seed in 2 3; do
timeout 1200 python tools/fuzz_parity.py 100 $seed 2>&1 | grep -v Warning > gpurun_out/fuzz$seed.log; grep '"ok": false' gpurun_out/fuzz$seed.log | cut -c1-700 | head -10; tail -1 gpurun_out/fuzz$seed.log
done
