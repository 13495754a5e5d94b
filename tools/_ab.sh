cd /root/repo
timeout 600 python -m pytest tests/test_gpu_grad.py -q -x -k "attention_bwd" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200
