cd "${GRAFT_REPO_ROOT:-/root/repo}"
M=gpu__time_duration.sum,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
for v in tools/vship/libmec_3f1eeb8.so minio_b200/libminio_ec.so; do
  MEC_LIB=$PWD/$v timeout 300 ncu --metrics $M --clock-control none -k regex:fused_rs_hh -c 6 --csv --log-file gpurun_out/cmp_$(basename $v).csv python tools/bench_configs.py jit > /dev/null 2>&1
  python - "$v" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/cmp_'+sys.argv[1].split('/')[-1]+'.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]: d.setdefault(r[ii],{'k':r[ki][:50]})[r[mi].split('__')[-1]]=r[vi]
print(sys.argv[1]); print(d[sorted(d,key=int)[-1]])
PY
done
