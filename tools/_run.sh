python -m pytest tests/test_search.py tests/test_neb.py tests/test_process_search.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-c3 --no-c4 --no-c5 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo rc=$?; tail -3 gpurun_out/r2b_bench.err
python - <<EOF2
import json
d=json.load(open("gpurun_out/r2b_bench.json"))
s=d["search"]
print({k:v for k,v in s.items() if k not in ("config","dimer","streamed")})
print("dimer", s["dimer"])
print("streamed", s["streamed"])
EOF2
