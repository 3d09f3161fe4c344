python bench.py > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
./tools/fp64_pipes > gpurun_out/r2_fp64_pipes.txt 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu --no-others --e2e-iters 50 > gpurun_out/r2_plain_small.json 2>gpurun_out/r2_plain_small.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-others --e2e-iters 50 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/prof_run.py c5 3 60 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:k_sweep_fast -s 126 -c 1 -o gpurun_out/r2_sweep_c5_e python tools/prof_run.py c5 3 60 > gpurun_out/r2_ncu_e.log 2>&1
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2_bench_c5.json")); r=j["roofline"]; o=j["other_workloads"]
print("value", j["value"], "ms", j["ms_per_step"], "frac", r["frac"], "launch", r["avg_launch_ms"], "param", r["param_kernel_ms_per_iter"])
print("e2e", j["e2e"]["value"], j["e2e"]["wall_s"], j["e2e"]["setup_ms"], "cpu", j["cpu_baseline"]["value"], "launches", j["gpu_launches"], j["clocks"])
print({k:(o[k].get("us_per_iteration"), o[k].get("hbm_frac_sweep_kernel"), o[k].get("spot_updates_per_s")) for k in ("c2","c4","c3")}); print(o["c3"]["adaptive"], o["c3"]["roofline"]["frac"]); print(o["qtune"]["wall_s"], o["qtune"]["oracle_check"])
PY
cut -c1-300 gpurun_out/r2_bench_reference_arm.json
