python scripts/kernel_times.py zipf255 1024 65536 2>&1 | grep -E "k_build|k_tree|k_find|k_decode |mib"
python scripts/kernel_times.py zipf255 1024 16384 2>&1 | grep -E "k_build|k_tree|k_find|k_decode |mib"
python scripts/kernel_times.py zipf255 1024 4096 2>&1 | grep -E "k_build|k_tree|k_find|k_decode |mib"
python scripts/kernel_times.py fibonacci 1024 65536 2>&1 | grep -E "k_build|k_tree|mib"
python scripts/ab_encode.py 1024 2>&1 | head -8
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4
