for shape in "0 24 128 128 32 64 fwd" "0 24 128 128 32 32 fwd" "0 72 128 128 16 64 fwd" "0 24 256 256 16 16 fwd" "0 24 64 64 64 64 fwd" "0 24 64 64 64 128 fwd" "0 24 128 128 32 64 dgrad"; do
  for h in 0 1; do
    echo -n "HALO=$h  "; EKL_RW_HALO=$h python tools/conv_one.py $shape 5 2>&1 | tail -1
  done
done
