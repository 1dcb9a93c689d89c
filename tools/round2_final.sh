# round-2 artefacts on the GPU box: the default bench line, the ncu launch list of the same command, one full capture of the lattice road's kernels
LAB_TREE_TIMING=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err; echo bench rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r03e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03e_ncu_launches.log 2>&1; echo launches rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pack|k_lat_" -c 12 -o gpurun_out/r03e_lattice_16383 python tools/run_one.py gather kruskal 16383 1 > gpurun_out/r03e_ncu_full.log 2>&1; echo full rc=$?
ls -la gpurun_out | grep r03e
