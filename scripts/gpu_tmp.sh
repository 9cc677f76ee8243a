timeout 600 python scripts/tf32_model_check.py 2>&1 | tail -6
echo "--- qkv TF32 too"
HEALSWIN_PRECISE_QKV=0 timeout 600 python scripts/tf32_model_check.py 2>&1 | grep "tf32:"
echo "--- bench: qkv fp32 (default)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
echo "--- bench: all TF32"
HEALSWIN_PRECISE_QKV=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
