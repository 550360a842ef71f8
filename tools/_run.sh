python tools/eam_bench.py clocks > gpurun_out/r2k.txt 2>&1; grep -v "^  ctas\|list decision\|density barrier\|exit  " gpurun_out/r2k.txt
python -m pytest tests/test_forces.py -m gpu -x -q 2>&1 | tail -2
