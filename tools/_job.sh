mkdir -p gpurun_out
timeout 600 python tools/parity_report.py > gpurun_out/r2_parity_report.json 2> gpurun_out/r2_parity_report.err; echo rc=$?; tail -3 gpurun_out/r2_parity_report.err; cat gpurun_out/r2_parity_report.json
