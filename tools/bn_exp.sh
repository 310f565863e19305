for shape in "0 24 4 4 640 512 fwd" "0 72 4 4 640 512 fwd" "0 24 4 4 1024 512 fwd" "0 24 4 4 2048 1024 fwd" "2 24 8 8 1024 2048 fwd" "0 24 4 4 640 512 dgrad" "2 24 16 16 256 512 fwd" "2 72 8 8 256 512 fwd"; do
  for bn in 16 32 64 128; do
    echo -n "BN=$bn  "; EKL_TC_BN=$bn python tools/conv_one.py $shape 5 2>&1 | tail -1
  done
done
