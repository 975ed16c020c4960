for v in old new old new; do
  cp scratch/ab/lib_$v.so efficient-nerf_b200/lib/libr2l_b200.so
  echo "== $v"; python scratch/time_nerf_mlp.py 2>&1 | grep "S="
done
