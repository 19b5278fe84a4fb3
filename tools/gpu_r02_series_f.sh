#!/bin/bash
# round 2: hybrid item policy (series items for cross tables / n_gauss > 16 / 200+ rows)
mkdir -p gpurun_out
timeout 600 python tools/bench_variants.py --tune "${TUNES:-}" > gpurun_out/variants_hybrid.jsonl 2> gpurun_out/variants_hybrid.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants_hybrid.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python tools/series_check.py > gpurun_out/series_check.jsonl 2> gpurun_out/series_check.err; echo "series_check rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/series_check.jsonl'):
    r = json.loads(line)
    print(r['shape'], 'cen', r['cen_max_abs_dev'], 'sat', r['sat_max_rel_dev'], r['sat_max_abs_dev_over_rowmax'], 'ngal', r['ngal_max_rel_dev'], 'xi', r['xi_max_dev_over_max_xi'], 'occ ms', round(r['occupation_ms_series'],3), round(r['occupation_ms_nodes'],3), 'pred ms', round(r['predict_ms_series'],3), round(r['predict_ms_nodes'],3))
PY
