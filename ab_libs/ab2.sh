mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/t_all3.log 2>&1; echo "all rc=$?" >> gpurun_out/t_all3.log
python profiles/experiments/check_tc.py > gpurun_out/check_tc.log 2>&1; echo "rc=$?" >> gpurun_out/check_tc.log
for rep in 1 2; do
for v in base new; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v burst rep $rep" >> gpurun_out/ab2.log
  SCMC_LIB=$lib STREAM=2 python profiles/experiments/ab_tma.py 8192 5 6 5 3 2>&1 | grep best >> gpurun_out/ab2.log
done; done
for v in base new base new; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v sustained" >> gpurun_out/ab2.log
  bash profiles/experiments/sustained.sh 6 $lib >> gpurun_out/ab2.log 2>&1
done
