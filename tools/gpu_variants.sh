for v in "CSM_L2_AHEAD_KB=0" "CSM_L2_AHEAD_KB=128" "CSM_L2_AHEAD_KB=512" "CSM_EVICT_FIRST=0" "CSM_L2_AHEAD_KB=0 CSM_EVICT_FIRST=0" "CSM_RING_SLOTS=3" ; do
  echo "=== $v"; env $v timeout 300 python tools/phase_profile.py --batch 1 2>&1 | head -8
done
