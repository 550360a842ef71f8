for t in 9 10 12 14 16; do python tools/pair_bench.py morse_pt 256 $t 2>&1 | tail -1 | cut -c1-150; done
for t in 12 16; do python tools/pair_bench.py morse_pt 256 $t 2 2>&1 | tail -1 | cut -c1-150; done
