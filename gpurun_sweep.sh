mkdir -p gpurun_out
ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --import-source on --clock-control none -k regex:k_walkers -c 8 -f -o gpurun_out/r1c_walkers python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1c_walkers.log 2>&1
tail -3 gpurun_out/r1c_walkers.log | cut -c1-300
