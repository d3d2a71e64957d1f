for m in 3 2 1; do echo "== REGTR_MHA_DEBUG=$m"; REGTR_MHA_DEBUG=$m CUDA_LAUNCH_BLOCKING=1 timeout 60 python scripts/test_mha_tc.py 2>&1 | grep -E "lens|error|Error|OK|FAIL" | head -4; done
