cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/prof_cm.py 512 30 > $O/r2x_plain_cm.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:pcg_fused -c 1 -o $O/r2x_pcg512 python tools/prof_cm.py 512 30 > $O/r2x_ncu_cm.log 2>&1
python tools/prof_k1.py 256 256 256 > $O/r2x_plain_k1.log 2>&1 && ncu --set full --clock-control none -k regex:"surface_scan|profile_kernel|fused_warp|bone_box|snr_reverb" -c 16 -o $O/r2x_k1k3_256 python tools/prof_k1.py 256 256 256 > $O/r2x_ncu_k1.log 2>&1
python tools/prof_k1.py 1024 512 512 > $O/r2x_plain_k1b.log 2>&1 && ncu --set full --clock-control none -k regex:"surface_scan|profile_kernel|fused_warp|bone_box|snr_reverb" -c 16 -o $O/r2x_k1k3_512 python tools/prof_k1.py 1024 512 512 > $O/r2x_ncu_k1b.log 2>&1
python bench.py --steps 2 --warmup 1 --no-stages --no-parity-check --no-cpu-cg --no-cpu-baseline --e2e-steps 0 > $O/r2x_bench_plain.json 2> $O/r2x_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2x_launches.csv python bench.py --steps 2 --warmup 1 --no-stages --no-parity-check --no-cpu-cg --no-cpu-baseline --e2e-steps 0 > $O/r2x_bench_ncu.log 2>&1
ls -la $O/r2x*
# the merge back is limited to 64 MiB: keep the raw pages of the K1 / K3 reports as CSV, drop the reports themselves
for r in r2x_k1k3_256 r2x_k1k3_512; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; rm -f $O/$r.ncu-rep; done
ls -la $O/r2x*
