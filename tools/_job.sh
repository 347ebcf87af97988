mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile > gpurun_out/r02d_plain.json 2> gpurun_out/r02d_plain.err; echo "plain rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-extra-configs --no-kernel-profile > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02d_launches.csv; ls -la gpurun_out/r02d_launches.csv
