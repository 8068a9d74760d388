timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "page_locked or host_lane or length_beyond" 2>&1 | tail -5
echo "== e2e probe, default (pin after 4)"
python scripts/e2e_probe.py 1024 reuse 2>&1 | grep -E "iter"
