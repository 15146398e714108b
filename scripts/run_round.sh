timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_win4 -s 6 -c 1 -o gpurun_out/prof_win4_c4 -f python scripts/probe_py.py c4 2 > gpurun_out/ncu_win4_c4.log 2>&1
tail -2 gpurun_out/ncu_win4_c4.log
