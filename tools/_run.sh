mkdir -p gpurun_out
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo bench rc=$?; tail -2 gpurun_out/r2n_bench.err
python bench.py --impl reference > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-c1 --no-c3 --no-c4 --no-c5 --no-search > gpurun_out/r2n_launches_bench.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_eam_n3 -s 6 -c 1 -o gpurun_out/r2n_n3 -f python tools/eam_bench.py steps=6 > gpurun_out/r2n_n3_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2n_n3.ncu-rep --page details > gpurun_out/r2n_n3_details.txt 2>/dev/null
ncu -i gpurun_out/r2n_n3.ncu-rep --page raw --csv > gpurun_out/r2n_n3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2n_n3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2n_n3_src.csv 2>/dev/null
python tools/eam_bench.py clocks > gpurun_out/r2n_phase_clocks.txt 2>&1
compute-sanitizer --tool memcheck python tools/eam_bench.py steps=1 2>&1 | tail -3 > gpurun_out/r2n_memcheck.txt
compute-sanitizer --tool racecheck python tools/eam_bench.py steps=1 2>&1 | tail -3 > gpurun_out/r2n_racecheck.txt
cat gpurun_out/r2n_memcheck.txt gpurun_out/r2n_racecheck.txt
