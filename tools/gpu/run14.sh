cd $GRAFT_REPO_ROOT
APP="python tools/stage_times.py --kind markov --size-mb 256 --no-stage --cpu-gen"
for spec in "k_scatter 2" "k_ent_cost 1" "k_ent_pm 1" "k_mtf_seq 0" "k_ent_sweep 1" "k_ranks 0" "k_rank_sort 3"; do set -- $spec
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:^$1\$ --launch-skip $2 --launch-count 1 -o gpurun_out/r2_$1 -f $APP > /dev/null 2> gpurun_out/ncu_$1.err
  ls -la gpurun_out/r2_$1.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
