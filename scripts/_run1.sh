python scripts/kernel_times.py zipf255 1024 65536 2>&1 | grep -E "k_tree|k_decode |mib"
python scripts/kernel_times.py zipf255 1024 4096 2>&1 | grep -E "k_tree|k_decode |mib"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "foreign or golden or c_api or matrix" 2>&1 | tail -3
