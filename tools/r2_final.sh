#!/bin/bash
# final evidence of round 2 on one GPU: every GPU test, both bench arms, launch lists and traces of the inflate paths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/r2_tests_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err
GZ_HOST=0 FQGPU_GZ_TRACE=1 timeout 600 python tools/gz_time.py 4000000 100 > gpurun_out/r2_gzip_trace.txt 2>&1
GZ_HOST=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_gzip_launches.csv python tools/gz_time.py 4000000 100 > gpurun_out/gz_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_bgzf_launches.csv python tools/bgzf_time.py 4000000 > gpurun_out/bgzf_ncu.log 2>&1
GZ_HOST=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gz_sync_kernel|gz_count_kernel|gz_write_kernel" -c 3 -f -o gpurun_out/r2_gzip_kernels python tools/gz_time.py 4000000 100 > gpurun_out/gz_ncu_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fq_meta_seg_kernel -s 12 -c 1 -f -o gpurun_out/r2_meta_seg python bench.py --steps 1 --warmup 1 --records 20000000 --meta-records 20000000 --no-ont --no-gz --no-ingest --no-e2e --no-cpu-baseline > gpurun_out/meta_ncu2.log 2>&1
python - <<'PY'
import csv, collections
for name in ('gzip', 'bgzf'):
    rows = [r for r in csv.reader(l for l in open(f'gpurun_out/r2_{name}_launches.csv') if l.startswith('"'))]
    h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        a = agg.setdefault(r[ki].split('(')[0], [0, 0.0]); a[0] += 1; a[1] += v
    with open(f'gpurun_out/r2_{name}_launches_summary.txt', 'w') as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; tools/{'gz' if name == 'gzip' else 'bgzf'}_time.py 4000000 (1.44 GB raw); all repetitions summed\n")
        for k, (n, t) in agg.items(): f.write(f"{k:44s} launches={n:4d} total={t/1e6:9.3f} ms\n")
    print(open(f'gpurun_out/r2_{name}_launches_summary.txt').read())
PY
tail -3 gpurun_out/r2_tests_gpu.log; cut -c1-300 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/bench_n1.err; tail -12 gpurun_out/r2_gzip_trace.txt
