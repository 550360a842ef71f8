for i in 1 2 3; do python bench.py --no-c1 --no-c3 --no-c4 --no-c5 --no-search --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['cpp_shim']['ms_per_step'])"; done
