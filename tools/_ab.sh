cd /root/repo
echo "token: $(timeout 120 python tools/attn_bwd_bench.py 16 2048 10 2>&1 | tail -1)"
echo "no token: $(LGB200_ATTN_BWD_TOKEN=0 timeout 120 python tools/attn_bwd_bench.py 16 2048 10 2>&1 | tail -1)"
echo "token 64x512: $(timeout 120 python tools/attn_bwd_bench.py 64 512 10 2>&1 | tail -1)"
echo "no token 64x512: $(LGB200_ATTN_BWD_TOKEN=0 timeout 120 python tools/attn_bwd_bench.py 64 512 10 2>&1 | tail -1)"
timeout 600 python -m pytest tests/test_gpu_grad.py -q -x 2>&1 | tail -2 | cut -c1-300
