cd /root/repo
LGB200_LIB=glue_factory_colon_b200/lib/var/a_msub0.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "attention" -x 2>&1 | tail -3
for lib in "" a_msub0 a_msub0_p4 ""; do
L=""; [ -n "$lib" ] && L=glue_factory_colon_b200/lib/var/$lib.so
echo -n "lib=$lib  "; LGB200_LIB=$L timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-library --no-fp32-mode 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],3), 'frac', round(r['frac'],4), 'alone', r.get('alone',{}).get('frac'), d['clocks']['sm_mhz'])
"
done
