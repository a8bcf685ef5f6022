for v in "CSM_REPL=1" "CSM_REPL=8" "CSM_REPL=32 CSM_L2_AHEAD_KB=0" "CSM_REPL=32 CSM_EVICT_FIRST=0" "CSM_REPL=32 CSM_L2_AHEAD_KB=0 CSM_EVICT_FIRST=0" "CSM_REPL=32 CSM_L2_AHEAD_KB=1024"; do
  echo "=== $v"; env $v timeout 300 python tools/phase_profile.py --batch 1 2>&1 | head -8
done
