cd /root/repo
for rc in 2049 2048 2045; do ASSIGN_RC=$rc timeout 120 python tools/assign_bench.py 2>&1 | grep -E "R = C|pass 2"; done
