python scripts/phase_prof.py run zipf255 1024 2>&1 | tail -14
for shape in zipf255 geometric fibonacci uniform english; do
  python scripts/kernel_times.py $shape 1024 65536 2>&1 | grep -E "k_decode |mib"
done
