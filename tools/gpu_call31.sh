#!/bin/bash
C=dynamic_llava_b200/csrc
cp $C/libdynllava_b200.so /tmp/lib_cur.so
for v in old new old new; do
  cp tools/ab/lib_$v.so $C/libdynllava_b200.so
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(r['value'],1), 'prefill', round(r['phases']['clip_prefill_ms'],2), 'decode', round(r['phases']['decode_ms_per_token_step'],3))"
done
cp /tmp/lib_cur.so $C/libdynllava_b200.so
