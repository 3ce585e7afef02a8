#!/bin/bash
# sustained.sh MODE [LIB]: the bench's device-resident figure (>= 1.2 s timed region, power cap in effect) with clocks, for A/B under sustained load
m=$1; lib=$2
[ -n "$lib" ] && export SCMC_LIB=$lib
python bench.py --mode $m --no-cpu-baseline --no-dense --extra-hits 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $m $lib value %.4e e2e %.4e' % (d['value'], d['e2e']['value']), d['clocks'])"
