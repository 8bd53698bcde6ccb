cd /root/repo
timeout 900 python -m pytest tests/test_device_linear_algebra.py tests/test_lm.py -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest_gpu.log
CB200_VERBOSE=1 CB200_SOLVER_TIMING=1 timeout 600 ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=6 > gpurun_out/r2h_solve_L.txt 2>&1; echo "solve rc=$?"; grep -E "cb200|iteration [0-9]|Linear solver|Minimizer  |Jacobian &|Residual only|Preprocessor" gpurun_out/r2h_solve_L.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2h_solve_launches.csv ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=1 > gpurun_out/r2h_solve_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2h_solve_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:40]:
    print(r[ki][:70].ljust(70), r[vi], r[ui])
PY
