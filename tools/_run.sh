EONB200_BENCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 > gpurun_out/r2z_bench_8gpu.out 2> gpurun_out/r2z_bench_8gpu.err; echo rc=$?
grep "\[bench\]" gpurun_out/r2z_bench_8gpu.err | cut -c1-60
python - <<EOF2
import json
txt=open("gpurun_out/r2z_bench_8gpu.out").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_sustained_400_calls"], d["e2e"]["value"])
s=d["search"]
print("search", s["searches_per_hour"], "queued", s["queued"]["wall_s"], s["queued"]["outcomes_differing_from_streamed"], "streamed", s["streamed"]["searches_per_hour"])
print("c3", d["extra"]["c3"]["ms_per_step"], d["extra"]["c3"]["search"]["ms"])
EOF2
