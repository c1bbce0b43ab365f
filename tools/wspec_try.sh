# usage: bash tools/wspec_try.sh   -- parity subset for the n = 4 kernels, then their timings
timeout -k 10 240 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "wspec or golden or sweep_envelope or unaligned or partially or stream" -x 2>&1 | tail -15
echo "wspec parity rc=$?"
timeout -k 10 300 python tools/quickbench.py --configs c5_f64,c5_f32,c3 2>&1 | tail -5
timeout -k 10 300 python tools/quickbench.py --configs c5_f64,c5_f32 --path wspec 2>&1 | tail -5
