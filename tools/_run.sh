python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo bench rc=$?; tail -2 gpurun_out/r2x_bench.err
python bench.py --impl reference > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_ref.err; echo ref rc=$?
python - <<EOF2
import json
d=json.load(open("gpurun_out/r2x_bench.json")); r=json.load(open("gpurun_out/r2x_bench_reference.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["cpp_shim"]["ms_per_step"], "ratio", d["e2e"]["value"]/r["value"], "launches", d["gpu_launches"], d["clocks"])
print(d["search"]["searches_per_hour"], d["search"]["mode"], d["search"]["queued"]["outcomes_differing_from_streamed"])
print(d["extra"]["c4"]["force_kernel"]["fp64_roofline_frac"], d["extra"]["c3"]["fp64_roofline_frac"])
EOF2
