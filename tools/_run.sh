python tools/pair_bench.py morse_pt 256 2>&1 | tail -1 | cut -c1-170
python tools/pair_bench.py morse_pt 1024 2>&1 | tail -1 | cut -c1-170
python tools/pair_bench.py morse_pt 32 2>&1 | tail -1 | cut -c1-170
python tools/pair_bench.py lj 4096 2>&1 | tail -1 | cut -c1-170
python -m pytest tests/test_forces.py tests/test_hessian.py tests/test_prefactor.py tests/test_process_search.py -m gpu -x -q 2>&1 | tail -1
