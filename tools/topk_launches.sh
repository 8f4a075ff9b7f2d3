#!/bin/bash
# usage: tools/topk_launches.sh <tag>   (on the GPU box) -- top-k parity tests, bench line, per-kernel launch times of one batch
tag=$1
timeout 300 python -m pytest tests/test_gpu_topk.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
MFB200_TOPK_STATS=1 timeout 200 python tools/bench_topk.py 75776 500000 128 100 3 > gpurun_out/${tag}_topk.log 2>&1; tail -2 gpurun_out/${tag}_topk.log | cut -c1-330
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_topk.py 37888 500000 128 100 1 > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:]: print(r[ki][:60], r[vi])
PY
