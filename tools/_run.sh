python tools/eam_bench.py clocks steps=40 > gpurun_out/r2y.txt 2>&1; grep -v "^  ctas\|list decision\|density barrier\|exit  " gpurun_out/r2y.txt
python -m pytest tests/test_forces.py -m gpu -x -q 2>&1 | tail -1
