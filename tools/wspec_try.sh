timeout 120 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "wspec or golden or sweep_envelope or unaligned or partially" -x 2>&1 | tail -15
echo "wspec parity rc=$?"
timeout 300 python tools/quickbench.py --configs c3,c5_f64,c5_f32 2>&1 | tail -5
