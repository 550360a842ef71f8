# scratch script of the last gpurun call (development aid): GPU tests, then the bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
