python -m pytest tests/test_forces.py tests/test_neb.py -m gpu -x -q 2>&1 | tail -1
python tools/eam_bench.py c3 steps=10 | head -1
python tools/eam_bench.py mode=1 steps=20 | head -1
