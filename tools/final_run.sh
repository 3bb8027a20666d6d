#!/bin/bash
# Everything behind profiles/r02_final_*: run on the GPU box as `gpurun -- 'bash tools/final_run.sh'`, then
# `python tools/summarize_profiles.py f` here turns the .ncu-rep files into the small tracked summaries.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench20.json 2> gpurun_out/f_bench20.err
python bench.py --steps 200 --warmup 10 --stagger > gpurun_out/f_bench200.json 2> gpurun_out/f_bench200.err
python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/f_bench1000.json 2> gpurun_out/f_bench1000.err
for c in car_gtg_pb haul_push_point haul_push_car; do python bench.py --config $c --steps 40 --no-cpu-baseline > gpurun_out/f_bench_$c.json 2> gpurun_out/f_bench_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python tools/probe_reset.py > gpurun_out/f_reset.json 2>&1
python tools/probe_rollout.py > gpurun_out/f_rollout.txt 2>&1
python tools/probe_reset_host.py > gpurun_out/f_reset_host.txt 2>&1
python bench_kernels.py > gpurun_out/f_microbench.jsonl 2> gpurun_out/f_microbench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_step_coop --launch-skip 604 --launch-count 1 -f -o gpurun_out/f_coop_late python tools/late_phase.py point_gtg 600 > gpurun_out/f_ncu_coop.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_step_free --launch-skip 604 --launch-count 1 -f -o gpurun_out/f_free_late python tools/late_phase.py point_gtg 600 > gpurun_out/f_ncu_free.log 2>&1
ncu --set full --clock-control none -k regex:k_reset --launch-skip 1 --launch-count 1 -f -o gpurun_out/f_reset python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_reset.log 2>&1
ls -la gpurun_out
