# round-end style verification on one GPU: full gpu test suite, smoke, both bench arms, ncu evidence for profiles/
python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_reference_r1.json 2> gpurun_out/bench_reference_r1.err; tail -1 gpurun_out/bench_reference_r1.json | cut -c1-400
python bench.py > gpurun_out/bench_ours_r1.json 2> gpurun_out/bench_ours_r1.err; tail -1 gpurun_out/bench_ours_r1.json | cut -c1-1800
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:row_move_vec -s 4 -c 2 -o gpurun_out/prof_gather_r1b -f python bench.py --steps 3 --warmup 3 --rows-per-gpu 20000000 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
