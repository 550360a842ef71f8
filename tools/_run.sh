python tools/pair_bench.py 2>&1 | tail -4
python tools/eam_bench.py c3 steps=5 2>&1 | head -1
python bench.py --steps 10 --warmup 3 --no-c1 --no-c3 --no-c4 --no-c5 --no-cpu > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo rc=$?
python - <<EOF2
import json
d=json.load(open("gpurun_out/r2l_bench.json"))
s=d["search"]
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"])
print({k:v for k,v in s.items() if k not in ("config","dimer","streamed")})
print("dimer", s["dimer"])
print("streamed", s["streamed"])
EOF2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
