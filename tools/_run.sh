for c in 1 2 3 4; do EONB200_PIPE_CHUNKS=$c timeout 120 python tools/e2e_py.py steps=300 2>&1 | head -1; done
EONB200_ZERO_COPY=2 timeout 120 python tools/e2e_py.py steps=300 2>&1 | head -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
EONB200_PIPE_CHUNKS=2 EONB200_TRACE=1 timeout 60 python tools/e2e_py.py steps=2 2>&1 | grep -v "^chunks" | sed -n 24,32p
