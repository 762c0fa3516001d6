#!/bin/bash
# prints one compact line per explore.py result
python tools/explore.py ${1:-all} ${2:-} ${3:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['label'], 'ms', d['ms_device'], 'ext_ms', d['ms_extend'], 'Mpaths/s', d['Mpaths_s'], 'Mseg/s', d['Mseg_s'], 'iters', d['iters'], 'n/seg', d['nodes_per_seg'], 'p/seg', d['prims_per_seg'])
"
