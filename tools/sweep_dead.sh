#!/bin/bash
# where / how wide the structure-blind UTF-8 pass runs (ETL_DEAD_MODE, ETL_DEAD_CTAS), c5 at full size
for mode in 3 2; do for ctas in 0; do
  echo "== mode $mode ctas $ctas"
  ETL_DEAD_MODE=$mode ETL_DEAD_CTAS=$ctas python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1),'GB/s', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})"
done; done
