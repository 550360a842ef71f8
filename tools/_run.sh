EONB200_TRACE=1 python tools/e2e_bench.py steps=2 chunks=1,4 errlines=10 2>&1 | tail -70
python tools/e2e_bench.py steps=50 chunks=1,2,4,8
