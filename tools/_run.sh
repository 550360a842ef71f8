for c in 2 3 4; do EONB200_PIPE_CHUNKS=$c timeout 120 python tools/e2e_py.py steps=300 2>&1 | head -1; done
timeout 120 python tools/e2e_py.py steps=300 2>&1 | head -1
python -m pytest tests/test_forces.py tests/test_host_shim.py -m gpu -x -q 2>&1 | tail -1
