python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo bench rc=$?; tail -2 gpurun_out/r2z_bench.err
python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_ref.err; echo ref rc=$?
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:k_eam_n3 -s 6 -c 1 -o gpurun_out/r2z_n3 -f python tools/eam_bench.py steps=6 > gpurun_out/r2z_n3_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2z_n3.ncu-rep --page details > gpurun_out/r2z_n3_details.txt 2>/dev/null
ncu -i gpurun_out/r2z_n3.ncu-rep --page raw --csv > gpurun_out/r2z_n3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2z_n3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2z_n3_src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-c1 --no-c3 --no-c4 --no-c5 --no-search > gpurun_out/r2z_launches_bench.log 2>&1; echo "launches rc=$?"
python tools/eam_bench.py clocks > gpurun_out/r2z_phase_clocks.txt 2>&1
python - <<EOF2
import json
d=json.load(open("gpurun_out/r2z_bench.json")); r=json.load(open("gpurun_out/r2z_bench_reference.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["cpp_shim"]["ms_per_step"], "ratio", d["e2e"]["value"]/r["value"])
print(d["search"]["searches_per_hour"], d["extra"]["c4"]["force_kernel"]["fp64_roofline_frac"], d["extra"]["c3"]["fp64_roofline_frac"], d["extra"]["c5"]["ms_per_band_evaluation_batch"])
EOF2
