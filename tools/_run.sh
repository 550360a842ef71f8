for r in 8 1 4; do echo "== rep $r"; EONB200_N3_REP=$r python tools/eam_bench.py clocks > gpurun_out/r2g_rep$r.txt 2>&1; grep -v "^  ctas\|list decision\|density barrier\|exit  " gpurun_out/r2g_rep$r.txt; done
echo "== rep 8 steps 100"; python tools/eam_bench.py steps=100 | head -1
