for c in 1 2 4; do EONB200_PIPE_CHUNKS=$c python tools/e2e_py.py steps=100 2>&1 | grep "pin 1"; done
