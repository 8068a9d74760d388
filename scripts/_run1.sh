for pad in 0 20000 60000 130000; do
  echo "== pad $pad"
  HUF_B200_DEC_PAD=$pad python scripts/kernel_times.py zipf255 1024 65536 2>&1 | grep -E "k_decode |zipf"
done
echo "== 4 KiB blocks"; python scripts/kernel_times.py zipf255 1024 4096 2>&1 | tail -20
echo "== 16 KiB blocks"; python scripts/kernel_times.py zipf255 1024 16384 2>&1 | tail -20
