cd /root/repo
CB200_SOLVER_TIMING=1 timeout 600 ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=4 > gpurun_out/r2g_solve_L.txt 2>&1; echo "solve rc=$?"; grep -E "iteration [0-9]|Linear solver|Minimizer  |Jacobian &|Residual only" gpurun_out/r2g_solve_L.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2g_solve_launches.csv ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=2 > gpurun_out/r2g_solve_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2g_solve_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:]:
    print(r[ki][:70].ljust(70), r[vi], r[ui])
PY
