python -m pytest tests/test_host_shim.py -m gpu -x -q 2>&1 | tail -4
EONB200_TRACE=1 tests/host/harness eon_b200/libeonb200.so 3 4 385 bench 3 pin < /dev/null 2>&1 | tail -3
