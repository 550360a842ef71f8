python -m pytest tests/test_search.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-c1 --no-c3 --no-c4 --no-c5 --no-cpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo rc=$?; tail -3 gpurun_out/r2q_bench.err
python - <<EOF2
import json
d=json.load(open("gpurun_out/r2q_bench.json"))
s=d["search"]
print({k:v for k,v in s.items() if k not in ("config","dimer","streamed","queued","one_launch","cpu_baseline")})
print("queued", s["queued"])
print("streamed", s["streamed"])
print("one_launch", s["one_launch"]["ms"], s["one_launch"]["searches_per_hour"])
EOF2
