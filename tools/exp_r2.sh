export LLKV_LEAN_ALLOW_R2=1
for cfg in 128,2,2,3 128,2,3,2 128,2,2,2 64,2,3,6 64,2,2,6 128,2,4,2; do echo "cfg $cfg"; python tools/prof_q.py 20000000 6 $cfg 2>&1 | grep "q1 [2345]" | awk '{printf "%s ", $3} END {print ""}'; done
