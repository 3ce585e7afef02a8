mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/t_final.log 2>&1; echo "all rc=$?" >> gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
for rc in 32 64 128; do echo "== rc $rc" >> gpurun_out/rc.log; SCMC_TMA_RC=$rc STREAM=2 python profiles/experiments/ab_tma.py 8192 5 6 2>&1 | grep best >> gpurun_out/rc.log; done
STREAM=2 python profiles/experiments/ab_tma.py 8192 1 6 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma_psi -s 8 -c 1 -o gpurun_out/r02_tc_final2 -f env STREAM=2 python profiles/experiments/ab_tma.py 8192 1 6 > gpurun_out/ncu_final2.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/b_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
