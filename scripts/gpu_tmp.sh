timeout 300 python scripts/tc_check.py 2>&1 | grep -E "BWD|FAIL"
