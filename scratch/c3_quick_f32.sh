#!/bin/bash
# quick C3 timing of the fp32 mode: prints step / e2e for a bench run with the given env
python bench.py --steps 3 --no-cpu --no-extra --dtype f32 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'step_ms', round(d['ms_per_step'],2), 'e2e', '%.3e' % d['e2e']['value'])"
