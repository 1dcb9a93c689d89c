# round-2 artefacts on the GPU box: the GPU suite, smoke(), the default bench line, the ncu launch list of the same command
# (a full ncu capture of the lattice road's kernels: see profiles/r02b_lattice_kernels_final_ncu.md for the command)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r03f_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r03f_smoke.log
LAB_TREE_TIMING=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err; echo bench rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r03f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03f_ncu_launches.log 2>&1; echo launches rc=$?
