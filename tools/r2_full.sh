#!/bin/bash
# the round-end sequence: every GPU test, the reference arm, the contract bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -25 > gpurun_out/tests_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -5 gpurun_out/tests_gpu.log; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('value','ms_per_step','stats_digest','gpu_launches','clocks')}
    keep['roofline'] = {k: d['roofline'][k] for k in ('achieved','frac','traffic')}
    for k in ('e2e','core_only','ont','gz','index','cpu_baseline'):
        v = d.get(k)
        keep[k] = ({kk: vv for kk, vv in v.items() if kk in ('value','frac_of_hbm_peak','roofline','error','seconds','fq_count_row','fq_meta','host_zlib','fq_meta_all_reads','chunks_inflated_on_device','false_block_starts_skipped')} if isinstance(v, dict) else v)
    keep['ingest'] = {k: (v.get('value') if isinstance(v, dict) else v) for k, v in (d.get('ingest') or {}).items()}
    print(json.dumps(keep, indent=1))
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/bench_n1.err
