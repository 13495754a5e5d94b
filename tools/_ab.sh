cd /root/repo
echo "base: $(timeout 120 python tools/attn_bench.py 2>&1 | tail -1)"
echo "attn3 event-driven: $(LGB200_ATTN3=1 timeout 120 python tools/attn_bench.py 2>&1 | tail -1)"
echo "attn3 lockstep issue: $(LGB200_ATTN3=1 LGB200_ATTN3_LOCKSTEP=1 timeout 120 python tools/attn_bench.py 2>&1 | tail -1)"
echo "base: $(timeout 120 python tools/attn_bench.py 2>&1 | tail -1)"
