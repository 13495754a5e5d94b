cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; tail -8 gpurun_out/gputest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
