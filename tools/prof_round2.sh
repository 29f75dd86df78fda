set -x
cd $GRAFT_REPO_ROOT
# (1) launch list of the headline leg
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --laps 0 --solve-laps 0 --gg-speeds 0 --single-laps 0 --no-cpu-baseline --cpu-solver-arms 0 > gpurun_out/r2_ncu_bench.log 2>&1
# (2) launch list from the middle of a 1024-lap solve (500-node laps)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 500 --csv --log-file gpurun_out/r2_solve1024_launches.csv python tools/ipm_batch.py 1024 500 > gpurun_out/r2_ncu_solve.log 2>&1
# (3) full capture of the warp-per-lap factorisation inside a 2048-lap solve, and of its solve kernel
ncu --set full --clock-control none --import-source on -k regex:k_kkt_factor_w -s 6 -c 1 -o gpurun_out/r2w_kkt_factor_w -f python tools/ipm_batch.py 2048 500 > gpurun_out/r2_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kkt_solve_w -s 12 -c 1 -o gpurun_out/r2w_kkt_solve_w -f python tools/ipm_batch.py 2048 500 > gpurun_out/r2_ncu_s.log 2>&1
# (4) full capture of the derivative kernel at 8000 nodes (traffic.json)
ncu --set full --clock-control none --import-source on -k regex:k_node_derivs -s 3 -c 1 -o gpurun_out/r2x_derivs -f python tools/prof_lap.py 8000 6 > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
