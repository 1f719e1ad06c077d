# DRAM traffic of one prediction-trunk launch at 4096 samples under the L2 residency policies (ncu cannot replay cooperative cluster launches: MZB_STACK_COOP=0)
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for cfg in "MZB_STACK_STICKY_MB=0" "MZB_STACK_STICKY_MB=45" "MZB_STACK_STICKY_MB=60" "MZB_STACK_STICKY_MB=105" "MZB_STACK_STICKY_MB=105 MZB_STACK_STICKY_FRAC=0.5" "MZB_STACK_STICKY_MB=45 MZB_STACK_OTHER_POLICY=1"; do
  echo "== $cfg"
  env $cfg MZB_STACK_COOP=0 ncu --metrics $M --clock-control none -k regex:conv_stack --csv python profiles/prof_stack.py 4096 2>/dev/null | grep -E "conv_stack" | awk -F'","' '{print $13, $15}' | tr -d '"' | paste - - - - | tail -2
done
