mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tma or tensor_core or tiled" > gpurun_out/t_tc.log 2>&1; echo "rc=$?" >> gpurun_out/t_tc.log
for rep in 1 2; do
for v in capc new rot iss3; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v burst rep $rep" >> gpurun_out/ab5.log
  SCMC_LIB=$lib STREAM=2 python profiles/experiments/ab_tma.py 8192 5 6 2>&1 | grep best >> gpurun_out/ab5.log
done; done
for v in capc new rot iss3 capc new; do
  if [ $v = new ]; then lib=semiclassicalmc.jl_b200/libscmc_b200.so; else lib=ab_libs/$v.so; fi
  echo "== $v sustained" >> gpurun_out/ab5.log
  bash profiles/experiments/sustained.sh 6 $lib >> gpurun_out/ab5.log 2>&1
done
