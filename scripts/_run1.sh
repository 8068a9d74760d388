for cfg in "0 0" "140 0" "144 0" "150 0" "136 1" "144 1"; do
  set -- $cfg
  echo "== fill $1 even $2"
  HUF_B200_DEC_FILL=$1 HUF_B200_DEC_EVEN=$2 python scripts/phase_prof.py run zipf255 1024 2>&1 | grep -E "k_decode|chunks "
done
