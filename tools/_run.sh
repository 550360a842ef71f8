python bench.py --steps 10 --warmup 3 --no-c1 --no-c3 --no-c5 --no-search --no-cpu > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo rc=$?; tail -3 gpurun_out/r2w_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['extra']['c4'])"
ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 1 -o gpurun_out/r2w_pair -f python tools/pair_bench.py > gpurun_out/r2w_pair_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2w_pair.ncu-rep --page details > gpurun_out/r2w_pair_details.txt 2>/dev/null
ncu -i gpurun_out/r2w_pair.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2w_pair_src.csv 2>/dev/null
