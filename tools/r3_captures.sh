cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/prof_cm.py 512 30 > $O/r3m_plain_cm.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:pcg_fused -c 1 -o $O/r3m_pcg512 python tools/prof_cm.py 512 30 > $O/r3m_ncu_cm.log 2>&1
python bench.py --steps 2 --warmup 1 --no-stages --no-parity-check --no-cpu-cg --no-cpu-baseline --e2e-steps 0 > $O/r3m_bench_plain.json 2> $O/r3m_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r3m_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --no-parity-check --no-cpu-cg --no-cpu-baseline --e2e-steps 0 > $O/r3m_bench_ncu.log 2>&1
ls -la $O/r3m*
