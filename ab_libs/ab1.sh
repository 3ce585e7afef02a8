mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/t_all2.log 2>&1; echo "all rc=$?" >> gpurun_out/t_all2.log
for rep in 1 2; do
for v in base new scalarprop gap8 gap16; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v burst rep $rep" >> gpurun_out/ab1.log
  SCMC_LIB=$lib STREAM=2 python profiles/experiments/ab_tma.py 8192 5 6 2>&1 | grep best >> gpurun_out/ab1.log
done; done
for v in base new scalarprop gap8 gap16 base new; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v sustained" >> gpurun_out/ab1.log
  bash profiles/experiments/sustained.sh 6 $lib >> gpurun_out/ab1.log 2>&1
done
