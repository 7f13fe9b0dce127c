export FAST_CHECK_ONLY_THROUGHPUT=1
for v in A B C D; do
  if [ $v = A ]; then unset RIMPHONY_B200_LIB; else export RIMPHONY_B200_LIB=$PWD/rimphony_b200/csrc/_exp$v/lib_$v.so; fi
  python tools/fast_check.py 0x3F 131072 > gpurun_out/abcd_$v.log 2>&1
  python tools/fast_check.py 0x3F 131072 > gpurun_out/abcd_${v}2.log 2>&1
  echo "variant $v: $(grep 'symphony only' gpurun_out/abcd_${v}2.log | cut -c1-170)"
done
