cd /root/repo
timeout 60 tools/micro/umma_rate 2>&1 | tee gpurun_out/umma_rate.log
