for v in "CSM_L2_AHEAD_KB=256" "CSM_L2_AHEAD_KB=0" "CSM_L2_AHEAD_KB=1024" "CSM_RING_SLOTS=3" "CSM_RING_SLOTS=4 CSM_L2_AHEAD_KB=512" ; do
  echo "=== $v"; env $v timeout 300 python tools/phase_profile.py --batch 1 2>&1 | head -8
done
